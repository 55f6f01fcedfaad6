"""Batched evaluation service with the contract of the reference's ``gpu_server.py`` (gpu_server.py:19-98).

Clients call ``network(feature_string)`` -- 196 int8 feature bytes in, ``(3332 float32 logit bytes, float value)`` out --
over msgpack-rpc on 127.0.0.1 (the wire protocol of the reference's ``mprpc``: requests ``[0, msgid, method, params]``,
responses ``[1, msgid, error, result]``).  Requests from all connections are marshalled into one batch of at most
``MARSHALL_COUNT`` positions or after ``MAXIMUM_WAIT_TIME`` seconds, whichever comes first (gpu_server.py:19-21,64-82),
and the batch is evaluated by ONE call of the GPU network (``az_net_forward_i8``).  Plain threads and sockets: gevent and
mprpc are not needed.

    python -m ataxxzero_b200.gpu_server model_path port
"""
import queue
import socket
import socketserver
import sys
import threading
import time

import msgpack
import numpy as np

FEATURE_BYTES = 7 * 7 * 4
POSTERIOR_BYTES = 7 * 7 * 17 * 4


class Processor(threading.Thread):
    """gpu_server.py:18-82: accumulate, evaluate together, hand every caller its row."""
    MAXIMUM_WAIT_TIME = 0.01
    MARSHALL_COUNT = 16

    def __init__(self, evaluate):
        """``evaluate(int8 features [B,7,7,4]) -> (float32 logits [B,7,7,17], float32 values [B,1])``"""
        super().__init__(daemon=True)
        self.evaluate = evaluate
        self.submit_queue = queue.Queue()
        self.batch_sizes = []
        self.stopping = False

    def submit(self, feature_string):
        if len(feature_string) != FEATURE_BYTES:
            raise ValueError("feature string must be %d bytes, got %d" % (FEATURE_BYTES, len(feature_string)))
        slot = queue.Queue(maxsize=1)
        self.submit_queue.put((bytes(feature_string), slot))
        result = slot.get()
        if isinstance(result, Exception):
            raise result
        return result

    def process(self, features, slots):
        if not slots:
            return
        self.batch_sizes.append(len(slots))
        try:
            batch = np.frombuffer(b"".join(features), dtype=np.int8).reshape(len(slots), 7, 7, 4)
            posteriors, values = self.evaluate(batch)
            posteriors = np.ascontiguousarray(posteriors, dtype=np.float32)
            for slot, posterior, value in zip(slots, posteriors, np.asarray(values, dtype=np.float32).reshape(-1)):
                slot.put((posterior.tobytes(), float(value)))
        except Exception as exc:                       # every waiting caller gets the failure, nobody hangs
            for slot in slots:
                slot.put(exc)

    def run(self):
        features, slots = [], []
        last_process_time = time.time()
        while not self.stopping:
            allowed = max(0.0, self.MAXIMUM_WAIT_TIME - (time.time() - last_process_time))
            try:
                feature_string, slot = self.submit_queue.get(timeout=allowed if slots else 0.05)
                if not slots:
                    last_process_time = time.time()   # the clock starts with the first request of a batch
                features.append(feature_string)
                slots.append(slot)
            except queue.Empty:
                pass
            if len(slots) >= self.MARSHALL_COUNT or (slots and time.time() - last_process_time >= self.MAXIMUM_WAIT_TIME):
                self.process(features, slots)
                features, slots = [], []
                last_process_time = time.time()


class _Handler(socketserver.BaseRequestHandler):
    def handle(self):
        unpacker = msgpack.Unpacker(raw=False)
        lock = threading.Lock()

        def answer(msgid, method, params):
            error, result = None, None
            try:
                if method != "network":
                    raise ValueError("unknown method %r" % (method,))
                posterior, value = self.server.processor.submit(params[0])
                result = [posterior, value]
            except Exception as exc:
                error = str(exc)
            with lock:
                self.request.sendall(msgpack.packb([1, msgid, error, result], use_bin_type=True))

        while True:
            data = self.request.recv(65536)
            if not data:
                return
            unpacker.feed(data)
            for message in unpacker:
                kind, msgid, method, params = message
                if kind != 0:
                    continue
                # one thread per in-flight call, so a single connection can pipeline requests into one batch
                threading.Thread(target=answer, args=(msgid, method, params), daemon=True).start()


class NetworkServer(socketserver.ThreadingTCPServer):
    allow_reuse_address = True
    daemon_threads = True

    def __init__(self, port, evaluate, host="127.0.0.1"):
        super().__init__((host, port), _Handler)
        self.processor = Processor(evaluate)
        self.processor.start()

    def shutdown(self):
        self.processor.stopping = True
        super().shutdown()


class RPCClient:
    """Minimal msgpack-rpc client with mprpc's ``call(method, *args)`` (what rpc_client.py:11-13 uses)."""

    def __init__(self, host, port):
        self.sock = socket.create_connection((host, port))
        self.unpacker = msgpack.Unpacker(raw=False)
        self.msgid = 0

    def call(self, method, *args):
        self.msgid += 1
        self.sock.sendall(msgpack.packb([0, self.msgid, method, list(args)], use_bin_type=True))
        while True:
            for kind, msgid, error, result in self.unpacker:
                if kind == 1 and msgid == self.msgid:
                    if error is not None:
                        raise RuntimeError(error)
                    return result
            data = self.sock.recv(1 << 16)
            if not data:
                raise ConnectionError("server closed the connection")
            self.unpacker.feed(data)

    def close(self):
        self.sock.close()


def gpu_evaluator(model_path, device=0, mode=None):
    """The real evaluator: the tensor-core network on `device`."""
    from . import Context, net
    ctx = Context(device=device)
    net.load_weights(ctx, model_path)
    mode = net.BF16 if mode is None else mode
    lock = threading.Lock()

    def evaluate(batch):
        with lock:
            return net.forward(ctx, batch, mode)
    return evaluate


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print("Usage: python -m ataxxzero_b200.gpu_server model_path port-to-host-on")
        return 1
    server = NetworkServer(int(argv[1]), gpu_evaluator(argv[0]))
    print("\nLaunching on port:", int(argv[1]))
    try:
        server.serve_forever()
    finally:
        server.server_close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

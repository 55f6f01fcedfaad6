"""Drop-in for the reference's ``link.py`` (link.py:6-32): the same four functions with the same
argument lists, bound to libataxxzero.so instead of ./cpp/self_play_client.so.

    launch_threads(output_path, visits, fill_buffer1, fill_buffer2, buffer_entries, thread_count)
    get_workload() -> 0 | 1
    complete_workload(workload, posteriors, values)
    shutdown()

The games run on the GPU; the caller is the evaluator, exactly as in
accelerated_generate_games.py:54-83.  (The reference's own argtypes list for complete_workload
omits the first int -- link.py:25-28 -- and only works because cdecl ignores it; the full
signature is declared here.)
"""
import ctypes

from . import _native

dll = _native.lib()
_native.register("launch_threads", None, [ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int])
_native.register("get_workload", ctypes.c_int, [])
_native.register("complete_workload", None, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p])
_native.register("shutdown", None, [])

launch_threads = dll.launch_threads
get_workload = dll.get_workload
complete_workload = dll.complete_workload
shutdown = dll.shutdown

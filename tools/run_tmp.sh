./tools/micro/int_alu_peak
python tools/profile_perft.py 7 > gpurun_out/perft_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_walk" -c 1 -o gpurun_out/perft_walk python tools/profile_perft.py 7 > gpurun_out/ncu_perft.log 2>&1
tail -2 gpurun_out/perft_plain.log gpurun_out/ncu_perft.log

SPARC_B200_LIB=build/lib_trd.so timeout 600 python tools/profile_amp.py --T 16 --launches 3 --batch 296 > gpurun_out/plain_trd.log 2>&1 &&
SPARC_B200_LIB=build/lib_trd.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/amp2_trd python tools/profile_amp.py --T 16 --launches 3 --batch 296 > gpurun_out/ncu_trd.log 2>&1
SPARC_B200_LIB=build/lib_tr.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp2_kernel -s 2 -c 1 -o gpurun_out/amp2_tr python tools/profile_amp.py --T 16 --launches 3 --batch 296 > gpurun_out/ncu_tr.log 2>&1
tail -n 2 gpurun_out/ncu_trd.log gpurun_out/ncu_tr.log

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2m
TDOA_DEMOD_VARIANT=23 timeout 300 python tools/diag_demod.py > gpurun_out/r2m/demod23.txt 2>&1; head -3 gpurun_out/r2m/demod23.txt;  sed -n 6p gpurun_out/r2m/demod23.txt
for v in 20 21 22 24 25 26; do TDOA_DEMOD_VARIANT=$v timeout 300 python tools/diag_demod.py --quick > gpurun_out/r2m/demod$v.txt 2>&1; echo $v; tail -n 1 gpurun_out/r2m/demod$v.txt; done
TDOA_DEMOD_VARIANT=23 ncu --set full --clock-control none --import-source on -k regex:k_demod_tma -c 1 -o gpurun_out/r2m/v23 python tools/diag_demod.py --quick > gpurun_out/r2m/ncu23.log 2>&1

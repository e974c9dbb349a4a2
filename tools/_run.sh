timeout 600 python -m pytest tests/test_gpu_msm.py -x -q -m gpu -k "affine" 2>&1 | tail -5
export BP_AFF_PASSES=3
bash tools/ncu_msm_pre.sh 0 20
ncu --set full --clock-control none --import-source on -k regex:k_aff_pass -s 2 -c 1 -f -o gpurun_out/aff_pass1 python tools/msm_probe.py --lgn 20 --iters 1 --c 16 --pre 0 > gpurun_out/ncu_aff.log 2>&1; tail -3 gpurun_out/ncu_aff.log

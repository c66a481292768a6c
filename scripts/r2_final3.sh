# round 2, last state: ncu --set full of the tensor-core nearest-neighbour kernel, then the final check (tests, smoke, bench,
# reference arm)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NNB_NN_ONE=1 ncu --set full --clock-control none --import-source on -k regex:nn_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_final_nn_tc \
    python scripts/dev/nn_time.py > gpurun_out/prof_r2_final_nn_tc.log 2>&1
ncu -i gpurun_out/prof_r2_final_nn_tc.ncu-rep --page raw --csv > gpurun_out/raw_r2_final_nn_tc.csv 2>/dev/null
bash scripts/final_check.sh

cd $GRAFT_REPO_ROOT
B="timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
NNB_TC_NPART=1 $B 2>/dev/null | python profiles/benchline.py generic_npart1
NNB_EXTRA_NVCC_FLAGS="-DNNB_TC_FIXED_D=30" python -m nnest_b200.build --force > /dev/null 2>&1
grep -A3 "mcmc_tc_kernelILi0ELi[12]" nnest_b200/lib/ptxas_nnb_tc.log | grep -i "spill\|registers"
$B 2>/dev/null | python profiles/benchline.py fixed30_npart2
NNB_TC_NPART=1 $B 2>/dev/null | python profiles/benchline.py fixed30_npart1
NNB_TC_NPART=1 $B --chains 16384 2>/dev/null | python profiles/benchline.py fixed30_npart1_16k
$B --chains 16384 2>/dev/null | python profiles/benchline.py fixed30_npart2_16k

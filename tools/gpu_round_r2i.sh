# K1 code shapes on 2^22 states, then one ncu --set full capture of the cluster tape evaluation (bundled tape) at 4096 proofs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/k1_variants.py > gpurun_out/k1_variants.json 2> gpurun_out/k1_variants.err; cat gpurun_out/k1_variants.json
bash tools/gpu_ncu_trace.sh k_tape_eval_cluster 4096 1
bash tools/gpu_ncu_trace.sh k_pair_tree_coop 4096 2

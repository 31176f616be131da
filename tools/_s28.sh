cd /root/repo
export ICP_DEBUG_CLUSTERS=1
echo "== correctness with clusters forced everywhere"
ICP_CLUSTER_ALL=1 timeout 600 python tools/gpu_bringup.py conv 2>&1 | grep -E "bring-up|max active|rror|CASE.*FAIL" | sort | uniq -c | head -20
ICP_CLUSTER_ALL=1 timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 5 2>&1 | tail -1
ICP_CLUSTER_ALL=1 timeout 300 python tools/gpu_bringup.py tower --precision fp16 --n 37 2>&1 | tail -1
echo "== conv4"
for CS in 2 8; do
ICP_CLUSTER_SIZE=$CS timeout 120 python tools/gpu_bringup.py layer --name conv4 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror|max active" | sed "s/^/cs=$CS /"
done
ICP_NO_CLUSTER=1 timeout 120 python tools/gpu_bringup.py layer --name conv4 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/none /"
echo "== other layers forced"
for L in gemm_k1024_n256 gemm_k1024_n192 conv3x3_k256_n256 c17_1x7_192; do
for CS in 2 8; do
ICP_CLUSTER_ALL=1 ICP_CLUSTER_SIZE=$CS timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/cs=$CS /"
done
ICP_NO_CLUSTER=1 timeout 120 python tools/gpu_bringup.py layer --name $L --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/none /"
done
echo "== tower"
for CS in 2 8; do
ICP_CLUSTER_SIZE=$CS timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 128 --graph --iters 20 --u8 2>&1 | grep SPEED | sed "s/^/cs=$CS /"
done
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['parity'])"

cd /root/repo
mkdir -p gpurun_out/r02
TL=tensorflow_web_deploy_b200/_ab/libinception_b200_tl.so
echo "== tall halo on 17x17 1x7"
timeout 120 python tools/gpu_bringup.py layer --name c17_1x7_192 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/im2col /"
ICP_HALO_TALL=1 timeout 120 python tools/gpu_bringup.py layer --name c17_1x7_192 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/tall   /"
ICP_HALO_TALL=1 ICP_HALO_T=4 timeout 120 python tools/gpu_bringup.py layer --name c17_1x7_192 --n 128 --precision fp16 --reps 20 2>&1 | grep -E "LAYER|rror" | sed "s/^/tall T4 /"
ICP_HALO_TALL=1 ICP_AB_LIB=$TL ICP_TIMELINE=gpurun_out/r02/tl_tall.bin timeout 120 python tools/gpu_bringup.py layer --name c17_1x7_192 --n 128 --precision fp16 --reps 2 > /dev/null 2>&1
python tools/timeline.py gpurun_out/r02/tl_tall.bin > gpurun_out/r02/tl_tall.txt 2>&1; head -48 gpurun_out/r02/tl_tall.txt
echo "== batch 1 kernel list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 85 --csv --log-file gpurun_out/r02/launches_b1.csv python tools/gpu_bringup.py speed --precision fp16 --n 1 --iters 2 --u8 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02/launches_b1.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=0
for r in rows[1:]:
    print(r[ki][:40], r[vi]); tot+=float(r[vi])
print('sum', tot)
PY
timeout 300 python tools/gpu_bringup.py speed --precision fp16 --n 1 --graph --iters 30 --u8 2>&1 | grep SPEED

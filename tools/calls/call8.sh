#!/bin/bash
# GPU call 8 of round 2 (8 GPUs): the box's PCIe / NUMA ceiling, configs[1] and configs[2] from one process with NUMA-local
# lanes, the config-5 stand-in with a bounded decoder thread count, the C++ segment scheduler on 8 lanes.
set -u
mkdir -p gpurun_out
O=gpurun_out
( nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" ) > $O/r02_c8_topo.txt 2>&1
timeout 300 python tools/pcie_ceiling.py 8 > $O/r02_c8_pcie_ceiling.txt 2>&1
timeout 400 python bench.py --single-process --gpus 8 --workload 720p_x4 --numa --no-cpu --steps 10 > $O/r02_c8_bench_single_n8_720p_numa.json 2> $O/r02_c8_bench_single_n8_720p_numa.err
timeout 400 python bench.py --single-process --gpus 8 --numa --no-cpu --steps 10 > $O/r02_c8_bench_single_n8_numa.json 2> $O/r02_c8_bench_single_n8_numa.err
timeout 400 python tools/bench_e2e.py --gpus 8 --frames 240 --numa > $O/r02_c8_e2e_n8_numa.txt 2>&1
REVE_DECODE_THREADS=1 timeout 400 python tools/bench_e2e.py --gpus 8 --frames 240 > $O/r02_c8_e2e_n8_dec1.txt 2>&1
T=/tmp/reve_seg
timeout 200 python tools/make_segments.py $T 16 12 1920 1080 2 > $O/r02_c8_sched.txt 2>&1
reve_b200/host/reve-upscale --segments $T -g 0,1,2,3,4,5,6,7 --random-weights -m /nonexistent \
    --encode-cmd 'ls {out_dir} | wc -l > {part}' >> $O/r02_c8_sched.txt 2>&1
echo "exit=$? parts: $(ls $T/video_parts | wc -l) frames in part 0: $(cat $T/video_parts/0.mp4) state: $(cat $T/video.temp | head -c 200)" >> $O/r02_c8_sched.txt
echo done

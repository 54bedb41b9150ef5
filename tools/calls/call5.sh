#!/bin/bash
# GPU call 5 of round 2 (8 GPUs): the product's shape -- ONE process, one thread + context per GPU -- on BASELINE.json
# configs[1] and configs[2], the config-5 stand-in with where-the-time-goes, and the C++ segment scheduler on 8 lanes.
set -u
mkdir -p gpurun_out
O=gpurun_out
( nproc; free -g | head -2; nvidia-smi --query-gpu=index,name,pcie.link.gen.current,pcie.link.width.current --format=csv ) > $O/r02_c5_box.txt 2>&1
timeout 400 python bench.py --single-process --gpus 8 --no-cpu > $O/r02_c5_bench_single_n8.json 2> $O/r02_c5_bench_single_n8.err
timeout 400 python bench.py --single-process --gpus 8 --workload 720p_x4 --no-cpu > $O/r02_c5_bench_single_n8_720p.json 2> $O/r02_c5_bench_single_n8_720p.err
timeout 400 python tools/bench_e2e.py --gpus 8 --frames 240 > $O/r02_c5_e2e_n8.txt 2>&1
T=/tmp/reve_seg
timeout 200 python tools/make_segments.py $T 16 12 1920 1080 2 > $O/r02_c5_sched.txt 2>&1
( /usr/bin/time -v reve_b200/host/reve-upscale --segments $T -g 0,1,2,3,4,5,6,7 --random-weights -m /nonexistent \
    --encode-cmd 'ls {out_dir} | wc -l > {part}' ) >> $O/r02_c5_sched.txt 2>&1
echo "exit=$? parts: $(ls $T/video_parts | wc -l) state: $(cat $T/video.temp | head -c 300)" >> $O/r02_c5_sched.txt
echo done

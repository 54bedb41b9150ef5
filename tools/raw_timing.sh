set -e
cd /root/repo
python - <<'PY'
import numpy as np
f = np.random.default_rng(0).integers(0,256,(1080,1920,3),dtype=np.uint8)
with open('/dev/shm/in.rgb','wb') as fo:
    for i in range(96): fo.write(f.tobytes())
PY
export REVE_HOST_TIMING=1
E=reve_b200/host/reve-upscale
( time $E --raw 1920x1080 -i /dev/shm/in.rgb -o /dev/shm/out.rgb -s 2 -m /nonexistent ) 2>&1 | grep -E "timing|real"
( time sh -c "cat /dev/shm/in.rgb | $E --raw 1920x1080 -i - -o - -s 2 -m /nonexistent > /dev/null" ) 2>&1 | grep -E "timing|real"
( time sh -c "cat /dev/shm/in.rgb | $E --raw 1920x1080 -i - -o - -s 2 -m /nonexistent | cat > /dev/null" ) 2>&1 | grep -E "timing|real"
rm -f /dev/shm/in.rgb /dev/shm/out.rgb

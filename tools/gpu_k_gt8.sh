# k = 9..16 at the C2 shape: the two-tile TMA pair against the CUDA-core kernels (us per update-iteration = device_ms / iters)
for k in 9 12 16; do
  for err in 1; do
    echo "k=$k two-tile TMA:"; python tools/profile_run.py --k $k --iters 40 --err $err | tail -1
    echo "k=$k CUDA-core:";    RESNMTF_TMA_GT8=0 python tools/profile_run.py --k $k --iters 40 --err $err | tail -1
  done
done
echo "k=8 TMA pair (reference point):"; python tools/profile_run.py --k 8 --iters 40 --err 1 --impl 3 | tail -1

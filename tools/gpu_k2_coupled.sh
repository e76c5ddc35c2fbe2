for p in 600 2500; do
for phi in 0 200; do
echo "p=$p phi=$phi default:"; python tools/profile_run.py --views 2 --phi $phi --n 40000 --p $p --k 5 --iters 60 --err 1 | tail -1
echo "p=$p phi=$phi two-pass:"; RESNMTF_IMPL=3 python tools/profile_run.py --views 2 --phi $phi --n 40000 --p $p --k 5 --iters 60 --err 1 | tail -1
done; done

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2v
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2v/pytest_multi.txt 2>&1; tail -n 4 gpurun_out/r2v/pytest_multi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --only-sharded > gpurun_out/r2v/sharded_n2.json 2> gpurun_out/r2v/sharded_n2.err; tail -c 300 gpurun_out/r2v/sharded_n2.err; cut -c1-700 gpurun_out/r2v/sharded_n2.json

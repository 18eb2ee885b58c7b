bash tools/tune_run.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('2gpu', d['value'], d['ms_per_step'], d['e2e'], d['extra']['stage_ms_rank0'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/prof_slab.py 2>&1 | tail -3
python -m pytest tests/test_slab_gpu.py -x -q -m gpu 2>&1 | tail -3

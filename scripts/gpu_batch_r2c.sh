# round-2 batch C (2 GPUs): where do the extra 0.16 ms per step at N > 1 come from?
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu --no-secondary --no-sustained"
echo "--- two independent single-GPU benches at once"
CUDA_VISIBLE_DEVICES=0 python bench.py $B > gpurun_out/c_ind0.json 2>gpurun_out/c_ind0.err &
P0=$!
CUDA_VISIBLE_DEVICES=1 python bench.py $B > gpurun_out/c_ind1.json 2>gpurun_out/c_ind1.err &
P1=$!
wait $P0 $P1
for f in c_ind0 c_ind1; do python -c "import json;d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]);print('$f',d['ms_per_step'],d['roofline']['kernel_ms'],d['e2e']['value'])"; done
run() { echo "--- torchrun n2 $*"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 $B "$@" 2>&1 | tail -1 > gpurun_out/$NAME.json; python -c "import json;d=json.load(open('gpurun_out/$NAME.json'));print(d['ms_per_step'],d['roofline']['kernel_ms'],d['per_rank'],d['e2e']['value'],d['e2e']['host_copy_ceiling'])"; }
PORT=29601 NAME=c_none run --exchange none
PORT=29602 NAME=c_peer run --exchange peer
PORT=29603 NAME=c_nccl run --exchange nccl
PORT=29604 NAME=c_peer200 run --exchange peer --steps 200
echo "--- torchrun n2 peer, python profile of the step loop (host time per step)"
B2S_HOST_TIMING=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29605 bench.py --gpus 2 $B --exchange peer 2>&1 | grep -i "host_step" | head -4

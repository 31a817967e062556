cd /root/repo
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" B200VIS_EXCHANGE_WHAT=slabs timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port $PORT tests/multi_gpu_parity.py --p2p > gpurun_out/p2p_$name.out 2> gpurun_out/p2p_$name.err
  echo "$name rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/p2p_$name.err | tail -12; tail -3 gpurun_out/p2p_$name.out
}
PORT=29701 run lean1 X=1
PORT=29702 run lean2 X=1
PORT=29703 run tma B200VIS_TILE_KERNEL=tma
PORT=29704 run lean3 X=1
timeout 200 python -m pytest tests/test_gpu_multi.py -q -x -k "one_process" 2>&1 | tail -3

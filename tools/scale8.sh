#!/bin/bash
# tools/scale8.sh: the end-to-end legs of bench.py on 1, 2, 4 and 8 GPUs of ONE box, then 8 GPUs again with the pinned
# buffers placed on the GPU's NUMA node (ARPEGGIO_NUMA_PIN=1, only with SCALE_NUMA=1); topology first.  Run under `gpurun --gpus 8`.
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"; } > gpurun_out/topology_r2.txt 2>&1
run() { # n tag env...
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then env "$@" python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --large-atoms 0 > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err
  else env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu --large-atoms 0 > gpurun_out/scale_${tag}.json 2> gpurun_out/scale_${tag}.err; fi
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
line = [l for l in open(f'gpurun_out/scale_{tag}.json') if l.startswith('{')]
if not line:
    print(tag, 'NO LINE'); sys.exit(0)
d = json.loads(line[-1])
e = d['e2e']
print('%-10s gpus %d | resident %.2e pairs/s | e2e %.2e (%.3f ms/step) serial %.2e records16 %.2e | pcie h2d %.1f d2h %.1f GB/s per GPU | batch %.0f structures/s' % (
    tag, d['n_gpus'], d['value'], e['value'], e['ms_per_step'], e['serial_value'], e['legs']['records16']['value'], e['pcie']['h2d_gbs'], e['pcie']['d2h_gbs'], d['batch']['value']))
PY
}
run 1 n1 A=1
run 2 n2 A=1
run 4 n4 A=1
run 8 n8 A=1
[ -n "$SCALE_NUMA" ] && run 8 n8_numa ARPEGGIO_NUMA_PIN=1

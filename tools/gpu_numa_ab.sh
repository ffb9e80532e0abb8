#!/bin/bash
# e2e (host buffers) with and without binding each rank to its GPU's NUMA node
N=${1:-1}
for numa in 1 0 1 0; do
  if [ "$N" = 1 ]; then
    HEHUB_B200_NUMA=$numa python bench.py --no-cpu --extras 0 2>/dev/null
  else
    HEHUB_B200_NUMA=$numa python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$numa bench.py --gpus $N --no-cpu --extras 0 2>/dev/null
  fi | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NUMA=$numa n_gpus', d['n_gpus'], 'value %.4e' % d['value'], 'e2e %.4e' % d['e2e']['value'], 'cpus bound', d['e2e'].get('host_cpus_bound_to_gpu_numa_node'))"
done
nvidia-smi topo -m 2>/dev/null | head -14

#!/bin/bash
# A/B of the CPU binding of multi-rank bench runs (e2e leg).  Usage: tools/e2e_bind_ab.sh N
N=${1:-8}
for mode in bind nobind; do
  if [ $mode = nobind ]; then export ML4CA_NO_BIND=1; else unset ML4CA_NO_BIND; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$((RANDOM % 10)) \
    bench.py --gpus $N --steps 30 --warmup 3 --skip-cpu --skip-extra > /tmp/ab_$mode.json 2>/dev/null
  python - "$mode" <<'PY'
import json, sys
d = json.loads(open("/tmp/ab_%s.json" % sys.argv[1]).read().splitlines()[-1])
print(sys.argv[1], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], d["e2e"].get("cpu_binding"))
PY
done

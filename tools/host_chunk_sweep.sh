#!/bin/bash
# e2e (host-buffer) env-steps/s of bench.py for several pipeline chunk sizes (ML4CA_HOST_CHUNK).  Tuning tool.
for c in 262144 524288 1048576 2097152 4194304; do
  ML4CA_HOST_CHUNK=$c python bench.py --steps 20 --warmup 3 --skip-cpu --skip-extra 2>/dev/null > /tmp/sweep_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open("/tmp/sweep_%s.json" % c).read().splitlines()[-1])
print(c, "e2e %.3f G env-steps/s" % (d["e2e"]["value"] / 1e9))
PY
done

#!/bin/bash
# Regenerates profiles/sass_r2.txt: opcode counts that show which objects carry tcgen05 / packed-FP32 / FP64 code.
cd "$(dirname "$0")/../ml4ca_b200/csrc" || exit 1
echo "# SASS opcode counts per object of libml4ca_b200.so (cuobjdump -sass, sm_100a; regenerate: tools/sass_counts.sh)"
echo "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, FFMA2/FMUL2/FADD2 = packed FP32, DFMA = FP64"
for o in policy.o ppo_update_tc.o ppo_update.o ppo_update_generic.o env_step_inst_4.o qp_alloc.o gae.o pinv_pid.o; do
  echo "== $o"
  cuobjdump -sass $o | grep -oE "\b(UTCHMMA|LDTM|STTM|UTCBAR|UTMALDG|UTMASTG|FFMA2|FMUL2|FADD2|HFMA2|DFMA|MUFU)\b" | sort | uniq -c | awk '{printf "%7d %s ", $1, $2} END {print ""}'
done

#!/bin/bash
# Where a one-shot `polypolish polish` process spends its wall time (POLYPOLISH_TIMING marks), small input.
cd "$(dirname "$0")/.."
D=$(mktemp -d)
python - "$D" <<'PY'
import sys
sys.path.insert(0, ".")
from polypolish_b200 import api
syn = api.Synth(seed=3, n_contigs=1, contig_len=200_000, depth=50.0)
syn.write(sys.argv[1])
PY
for i in 1 2 3; do
  s=$(date +%s.%N)
  POLYPOLISH_TIMING=1 build/polypolish polish --quiet "$D/draft.fasta" "$D/reads_1.sam" "$D/reads_2.sam" 2>&1 >/dev/null | grep -E "timing|Error"
  e=$(date +%s.%N)
  python -c "print(\"process wall %.3f s\" % ($e - $s))"
done
rm -rf "$D"

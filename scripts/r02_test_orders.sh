#!/bin/bash
# order-dependence check of the GPU suite: reversed order, and two shuffles (static launcher state, counter sets, graphs)
python -m pytest tests -m gpu -q --collect-only 2>/dev/null | grep "::" > /tmp/ids.txt
wc -l < /tmp/ids.txt
tac /tmp/ids.txt > /tmp/ids_rev.txt
python -m pytest -q -p no:cacheprovider $(cat /tmp/ids_rev.txt | tr '\n' ' ') 2>&1 | tail -4 | cut -c1-300
for seed in 1 2; do
  python - <<PY > /tmp/ids_shuf.txt
import random
ids = [l.strip() for l in open('/tmp/ids.txt')]
random.Random($seed).shuffle(ids)
print(' '.join(ids))
PY
  python -m pytest -q -p no:cacheprovider $(cat /tmp/ids_shuf.txt) 2>&1 | tail -4 | cut -c1-300
done

# dealing run length (chunks a CTA takes in a row) on the mixed-size and the small-chunk configs, tuning on
for a in "--workload config4" "--workload config5 --seconds 0.25"; do
  for cb in 2 4 8 16 32 64; do
    OHP_CHUNK_BLOCK=$cb python bench.py $a --no-cpu-baseline --no-e2e --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-32s chunk_block=$cb cap=%d frac %.3f' % (d['config']['workload'][:32], d['config']['inflight_chunks_per_cta'], d['roofline']['frac']))"
  done
done

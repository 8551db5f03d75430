for a in "--workload config2" "--workload config5 --seconds 0.25" "--workload config4" "--workload config3" "--workload mixed"; do
  for mode in 0 1; do for cap in 8 12 18 24 32; do
    OHP_SERIAL_PLACE=$mode OHP_CAP_CHUNKS=$cap python bench.py $a --no-cpu-baseline --no-e2e --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-32s serial=$mode cap=$cap frac %.3f' % (d['config']['workload'][:32], d['roofline']['frac']))"
  done; done
done

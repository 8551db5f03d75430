# pinned in-flight caps per placement mode (one library per mode: build/libohp_mode{0,1}.so), then the tuned run
for a in "--workload config2" "--workload config5 --seconds 0.25" "--workload config4" "--workload config3"; do
  for mode in 0 1; do for cap in 12 24 32; do
    OHP_LIB_CUDA=$PWD/build/libohp_mode$mode.so OHP_CAP_CHUNKS=$cap python bench.py $a --no-cpu-baseline --no-e2e --steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-32s mode=$mode cap=$cap frac %.3f' % (d['config']['workload'][:32], d['roofline']['frac']))"
  done; done
done
for a in "--workload config2" "--workload config5 --seconds 0.25" "--workload config5" "--workload config4" "--workload config3" "--workload mixed"; do
  for mode in 0 1; do
    OHP_LIB_CUDA=$PWD/build/libohp_mode$mode.so python bench.py $a --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('%-32s mode=$mode tuned cap=%d frac %.3f' % (d['config']['workload'][:32], d['config']['inflight_chunks_per_cta'], d['roofline']['frac']))"
  done
done

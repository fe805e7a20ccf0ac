# quick A/B helper: GPU parity tests of the decoder path + light bench line of every config
mkdir -p gpurun_out/exp
timeout 900 python -m pytest tests/test_gpu_nice.py tests/test_gpu_fullsize.py tests/test_gpu_mapping_iteration.py tests/test_gpu_tc.py -q -x 2>&1 | tail -8
for cfg in mapping tracking dense mesh256; do
  timeout 200 python bench.py --light --config $cfg 2> gpurun_out/exp/$cfg.err | tail -1 > gpurun_out/exp/$cfg.json
  python -c "
import json; d=json.load(open('gpurun_out/exp/$cfg.json')); print('$cfg', d['ms_per_step'], d['value'], d['roofline']['kernel_ms'])"
done

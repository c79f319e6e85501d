#!/bin/bash
# Final round-2 profile session (one B200): GPU tests, bench + reference arm, ncu launch list of a bench step,
# full-set captures of the head kernels, the batched leg and the batched projection.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
python -m overlapnet_b200.build > gpurun_out/r2_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log; grep "\[parity\]" gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
echo "bench exit $?"; tail -3 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "ref exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2_bench_ncu.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_delta_conv1_tc|k_conv2_sw_tc|k_conv3_pair_tc|k_corr_tc' -s 10 -c 4 \
  -o gpurun_out/r2_prof_heads python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/r2_ncu_heads.log 2>&1
echo "ncu heads exit $?"
cat > /tmp/one_leg.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch, bench
from overlapnet_b200 import synth
from overlapnet_b200.engine import Engine
eng = Engine(model=bench.MODEL, precision='f16_tc', max_batch_scans=64, max_batch_pairs=1)
eng.load_weights(bench.make_weights(4))
clouds = [synth.kitti_like_cloud(s) for s in range(32)]
b = eng.upload_clouds(clouds)
x = torch.from_numpy(synth.range_like_images(5, 8, 4)).to(eng.device).repeat(8, 1, 1, 1)
for _ in range(3):
  eng.preprocess(b)
  eng.leg(x)
torch.cuda.synchronize()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_leg_batched_tc|k_leg_layer1_direct|k_project_scatter|k_project_gather' -s 26 -c 13 \
  -o gpurun_out/r2_prof_leg_proj python /tmp/one_leg.py > gpurun_out/r2_ncu_leg.log 2>&1
echo "ncu leg/proj exit $?"
timeout 300 python tools/time_leg.py > gpurun_out/r2_time_leg.log 2>&1; cat gpurun_out/r2_time_leg.log
ls -la gpurun_out | grep r2_ | tail -20

timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_full_gpu_tests.log 2>&1; tail -2 gpurun_out/r2_full_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -c 400; echo
timeout 500 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["variants"]["gptq_int4_g128_llama3_8b"]["s_per_step"], d["variants"]["cfg2a_no_mse_clip0.9"]["roofline"]["frac"])
PY

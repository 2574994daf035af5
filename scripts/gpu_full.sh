# full GPU suite + benches (both families) + other configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_auto.json 2> gpurun_out/bench_auto.err
timeout 600 python bench.py --steps 3 --warmup 3 --kernel stream --no-cpu-baseline > gpurun_out/bench_stream.json 2> gpurun_out/bench_stream.err
timeout 1500 python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
tail -3 gpurun_out/pytest.log; tail -1 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_auto.json", "gpurun_out/bench_stream.json"):
    try:
        d = json.load(open(f)); r = d["roofline"]
        print(f, "value %.3e" % d["value"], "ms/step %.2f" % d["ms_per_step"], "frac %.3f" % r["frac"], "kernel_ms %.2f" % r["kernel_ms"],
              "e2e %.3e" % d["e2e"]["value"], "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-500:])
PY
cut -c1-260 gpurun_out/configs.jsonl; tail -2 gpurun_out/configs.err

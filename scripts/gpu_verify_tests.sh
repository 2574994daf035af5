mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/verify_smoke.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/verify_pytest.log; cat gpurun_out/verify_pytest.log

# final validation + evidence in one call: tests, smoke, benches, other configs, ncu captures
bash scripts/gpu_full.sh
bash scripts/gpu_profile.sh r1f > gpurun_out/profile_r1f.log 2>&1
tail -6 gpurun_out/profile_r1f.log

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
bash scripts/profile_round.sh r1k > gpurun_out/r1k_profile_round.log 2>&1; tail -5 gpurun_out/r1k_profile_round.log

for r in 0 1 0 1; do echo "realloc $r"; KC_TRUNK_REALLOC=$r python tests/diag_perf.py 2>&1 | sed -n 3,4p; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "half_batch or forward_matches or search_with_net" 2>&1 | tail -2
python tests/diag_search.py 160 graph

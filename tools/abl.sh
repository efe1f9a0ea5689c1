timeout 600 python -m pytest tests/test_gpu_ifnet.py tests/test_gpu_parity_r2.py -q -x -p no:cacheprovider 2>&1 | grep -E "^E|passed|failed" | head -12
for nb in 3 2 4; do echo "nb=$nb"; SVR_FQ_NB=$nb timeout 300 python tools/fq_time.py 2>&1 | grep "halo=True" | cut -c1-90; done

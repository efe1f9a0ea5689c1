for d in 0 2 4 8 12 14; do echo "debug=$d"; SVR_FQ_DEBUG=$d python tools/fq_time.py 2>&1 | grep "halo=True" | cut -c1-80; done

for c in 4 3; do for ps in 3 4 5; do echo "c=$c per_sm=$ps"; HFB200_NTT_C=$c HFB200_NTT_PER_SM=$ps python tools/bench_kernels.py 2>&1 | head -1; done; done

for bal in 0 1; do for ord in 0 1 2 3; do
  echo "== BAL=$bal ORD=$ord"
  AACLIP_ATTN_BAL=$bal AACLIP_ATTN_ORD=$ord timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention 2>&1 | tail -1
  AACLIP_ATTN_BAL=$bal AACLIP_ATTN_ORD=$ord timeout 120 python tools/attn_probe.py 64
done; done

"""TEST INFRASTRUCTURE: CPU oracle + reference-CUDA harness. Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / reference arm may import this package."""

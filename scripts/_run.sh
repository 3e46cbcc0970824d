PPBO_OVERLAP_RESERVE=16 python scripts/timeline.py --mode steady --steps 2 --out gpurun_out/timeline_steady.txt --dump gpurun_out/seq_steady.txt 2>&1 | tail -1

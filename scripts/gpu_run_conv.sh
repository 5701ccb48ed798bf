#!/bin/bash
mkdir -p gpurun_out
RB200_TWO_TRACKS=0 timeout 300 python scripts/conv_times.py 2>&1 | grep tiled | tee gpurun_out/conv_one_track.log
RB200_TWO_TRACKS=1 timeout 300 python scripts/conv_times.py 2>&1 | grep tiled | tee gpurun_out/conv_two_tracks.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiled or conv_w_ref or batch_sizes or concurrent" 2>&1 | tail -3

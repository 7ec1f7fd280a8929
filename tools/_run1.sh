timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
CES_SWEEP_SHAPES="16384,64,256;65536,64,256;16384,64,50;16384,1024,4096" timeout 600 python tools/sweep.py 2>&1 | tail -5

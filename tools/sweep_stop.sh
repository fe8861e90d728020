python tools/ramp.py 6.0 2>&1 | tail -1

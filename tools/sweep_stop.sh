for f in 0.99 0.5; do C4_STOP_FRAC=$f python tools/ramp.py 6.0 2>&1 | tail -1; done

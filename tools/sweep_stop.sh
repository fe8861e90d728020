for v in "$@"; do env $v python tools/ramp.py 5.0 2>&1 | tail -1; done

for r in 1 4 8 16 24 32; do for st in 1 2 4; do echo -n "refill=$r steps=$st: "; PYR_TRACE_REFILL=$r PYR_TRACE_STEPS=$st python tools/profile_step.py 4 | grep -o "trace [0-9.]* ms"; done; done

# Sweep of the traversal kernel's lane-refill threshold and node steps per leaf step (dead ends).
for r in 6 8 12; do for st in 1 2 3 4; do echo -n "refill=$r steps=$st: "; PYR_TRACE_REFILL=$r PYR_TRACE_STEPS=$st python tools/profile_step.py 8 | grep -o "trace [0-9.]* ms"; done; done

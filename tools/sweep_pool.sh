# Throughput of the C2 step against the number of paths in flight ("Paths in flight").
for p in 131072 262144 524288 1048576 2097152; do echo -n "pool=$p: "; PYR_POOL=$p python tools/profile_step.py 8; done

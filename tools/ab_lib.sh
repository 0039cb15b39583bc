# usage: tools/ab_lib.sh "<profile_step args>" lib1.so lib2.so ...   (A/B of built library variants)
args="$1"; shift
cp pyrite_b200/libpyrite_b200.so /tmp/keep.so
for round in 1 2; do for lib in "$@"; do
  cp "$lib" /tmp/variant.so; cp /tmp/variant.so pyrite_b200/libpyrite_b200.so
  echo -n "$lib: "; python tools/profile_step.py $args
done; done
cp /tmp/keep.so pyrite_b200/libpyrite_b200.so

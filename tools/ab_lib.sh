# usage: tools/ab_lib.sh lib1.so lib2.so ...   (A/B of built library variants on the C2 step)
for lib in "$@"; do
  cp pyrite_b200/libpyrite_b200.so /tmp/keep.so
  cp "$lib" /tmp/variant.so; cp /tmp/variant.so pyrite_b200/libpyrite_b200.so
  echo -n "$lib: "; python tools/profile_step.py 8
  cp /tmp/keep.so pyrite_b200/libpyrite_b200.so
done

#!/bin/bash
# On the GPU box: time every build/variants/libikb200_<name>.so on the humanoid config (see build_humanoid_variants.sh).
cd "$(dirname "$0")/.."
cp ik_b200/libikb200.so /tmp/libikb200_default.so
for f in build/variants/libikb200_*.so; do
    name=$(basename $f .so); name=${name#libikb200_}
    cp $f ik_b200/libikb200.so
    for rep in 1 2; do echo -n "$name: "; python tools/humanoid_one.py ${1:-262144} 2>&1 | tail -1; done
done
cp /tmp/libikb200_default.so ik_b200/libikb200.so

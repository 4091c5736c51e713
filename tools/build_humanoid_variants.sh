#!/bin/bash
# A/B builds of the humanoid specialisation (rolled left-looking loops, block width) for one gpurun call:
#   build/variants/libikb200_<name>.so, swapped over ik_b200/libikb200.so by tools/run_humanoid_variants.sh
set -e
cd "$(dirname "$0")/.."
SPEC=ik_b200/specs/humanoid_limbs.json
cp $SPEC /tmp/humanoid_spec_orig.json
mkdir -p build/variants
for v in "r0w2:0:2" "r2w2:2:2" "r1w2:1:2" "r2w3:2:3" "r4w2:4:2"; do
    IFS=: read name rolled width <<< "$v"
    python3 - "$rolled" "$width" <<'PY'
import json, sys
p = "ik_b200/specs/humanoid_limbs.json"
s = json.load(open(p))
s["rolled_update"] = int(sys.argv[1])
s["parallel_block_width"] = int(sys.argv[2])
json.dump(s, open(p, "w"), indent=2)
PY
    make -j8 > /dev/null 2>&1
    cp ik_b200/libikb200.so build/variants/libikb200_$name.so
    echo "$name: $(grep -A2 'EdLi1ELi1ELb0' build/obj/spec_humanoid_limbs.ptxas.log | grep -o 'Used [0-9]* registers') $(grep -A1 'EdLi1ELi1ELb0' build/obj/spec_humanoid_limbs.ptxas.log | grep -o '[0-9]* bytes spill stores')"
done
cp /tmp/humanoid_spec_orig.json $SPEC
make -j8 > /dev/null 2>&1

#!/bin/bash
# static SASS instruction count + opcode classes of the nobel-eu first-fit step kernel in a built library
so=${1:-optical_networking_gym_b200/libqrmsa_b200.so}
fn=${2:-_ZN5qrmsa13k_step_policyILi320ELi6ELi5ELi0EEEvNS_7KParamsEi}
cuobjdump -sass -fun "$fn" "$so" | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/sass_fn.txt
echo "instructions: $(wc -l < /tmp/sass_fn.txt)"
awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); c[op]++} END {for (o in c) print c[o], o}' /tmp/sass_fn.txt | sort -rn | head -24 | tr '\n' ';'
echo

#!/usr/bin/env bash
# tools/sass_fn.sh <object> <substring of the mangled kernel name>: prints "addr instruction" lines of one kernel
set -euo pipefail
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" | grep -oE "Function : \S+" | awk '{print $3}' | grep -- "$pat" | head -1)
cuobjdump -sass -fun "$fn" "$obj" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s+\/\*.*$//'

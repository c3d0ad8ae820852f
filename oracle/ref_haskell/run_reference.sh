#!/bin/bash
# Run the unmodified reference on the golden cases (needs `stack`; see README.md in this directory).
#   oracle/ref_haskell/run_reference.sh /path/to/qubism-checkout
set -euo pipefail
here=$(cd "$(dirname "$0")" && pwd)
root=$(cd "$here/../.." && pwd)
ref=${1:?usage: run_reference.sh /path/to/qubitrot-qubism-checkout}
out="$root/oracle/_ref"
mkdir -p "$out"
python "$here/make_input.py" "$out"
cp "$here/Golden.hs" "$out/Golden.hs"
# compiled inside the reference's own project: its modules, its resolver (lts-12.4), its hmatrix
(cd "$ref" && stack build && stack ghc -- -O1 -isrc -outputdir "$out/obj" -o "$out/golden_ref" "$out/Golden.hs")
"$out/golden_ref" "$out/ref_input.txt" > "$out/ref_output.txt"
python "$here/collect.py" "$out/ref_output.txt"
echo "tests/golden/reference.json written: python -m pytest tests/test_oracle.py -q now pins the oracle"

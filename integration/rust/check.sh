#!/bin/bash
# Applies the CUDA overlay to a checkout of eternal-flame-AD/yume-pdq and runs `cargo check --features cuda`.
# usage: integration/rust/check.sh /path/to/yume-pdq-checkout
# Without cargo (this build image has none) it only checks that build_cuda.rs lists every CUDA source, prints SKIPPED
# and exits 0 -- the Rust glue has never been compiled; the C ABI it binds is what the test-suite exercises.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/../.." && pwd)
want=$(python3 - "$REPO" <<'PY'
import re, sys
src = open(sys.argv[1] + "/yume_pdq_b200/build.py").read()
print(" ".join(sorted(re.findall(r'"(\w+\.cu)"', re.search(r"SOURCES = \[(.*?)\]", src, re.S).group(1)))))
PY
)
have=$(grep -o '"[a-z0-9_]*\.cu"' "$HERE/build_cuda.rs" | tr -d '"' | sort | tr '\n' ' ' | sed 's/ $//')
if [ "$want" != "$have" ]; then
    echo "build_cuda.rs is stale: lists [$have], yume_pdq_b200/build.py builds [$want]" >&2
    exit 1
fi
echo "build_cuda.rs lists all $(echo $want | wc -w) CUDA sources"
if ! command -v cargo >/dev/null 2>&1; then
    echo "SKIPPED: no cargo in PATH (the overlay was not compiled)"
    exit 0
fi
CHECKOUT=${1:?path to a yume-pdq checkout}
WORK=$(mktemp -d)
cp -r "$CHECKOUT"/. "$WORK"/
mkdir -p "$WORK/cuda"
cp -r "$REPO/yume_pdq_b200/csrc" "$WORK/cuda/csrc"
cp -r "$REPO/include" "$WORK/cuda/include"
cp "$HERE/src/cuda_sys.rs" "$WORK/src/cuda_sys.rs"
cp "$HERE/src/kernel/cuda.rs" "$WORK/src/kernel/cuda.rs"
cp "$HERE/src/matching/cuda.rs" "$WORK/src/matching/cuda.rs"
cp "$HERE/examples/cuda.rs" "$WORK/examples/cuda.rs"
(cd "$WORK" && patch -p1 < "$HERE/Cargo.toml.patch")
cat "$HERE/build_cuda.rs" >> "$WORK/build.rs"
sed -i 's/^fn main() {/fn main() {\n    #[cfg(feature = "cuda")]\n    build_cuda();/' "$WORK/build.rs"
echo '#[cfg(feature = "cuda")] pub mod cuda_sys;' >> "$WORK/src/lib.rs"
echo '#[cfg(feature = "cuda")] pub mod cuda;' >> "$WORK/src/kernel/mod.rs"
echo '#[cfg(feature = "cuda")] pub mod cuda;' >> "$WORK/src/matching/mod.rs"
(cd "$WORK" && cargo check --features cuda --examples)

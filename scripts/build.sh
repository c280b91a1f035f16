#!/bin/bash
# Build the in-tree extension under the lock the GPU retry loop honors.
cd "$(dirname "$0")/.."
touch /tmp/psd_build.lock
python -c "
import os
os.environ.setdefault('TORCH_CUDA_ARCH_LIST','10.0')
from poseidon_b200.ops import build as b
b.build_extension(verbose=True)
" 2>&1 | grep -E "error|Error|FAILED|warning #|bytes spill" | grep -v " 0 bytes spill stores, 0 bytes spill loads" | head -40
rm -f /tmp/psd_build.lock
ls -la poseidon_b200/_ext/poseidon_b200_C.so

#!/bin/bash
# Vendor the UNMODIFIED reference PHet class into the git-ignored oracle/_ref/ so that it travels to the GPU box
# with the gpurun snapshot (nothing under /root/reference exists there) and `bench.py --impl reference` can time
# the reference's own code (cpu_baseline.kind = "reference") under oracle/ref_shim.py.  Build container only;
# nothing is copied into the tracked tree.  phet.py is self-contained: its other imports are third-party.
set -e
SRC=${PHET_REFERENCE_SRC:-/root/reference/src}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$SRC/model/phet.py" ] || { echo "reference tree not present at $SRC" >&2; exit 0; }
mkdir -p "$HERE/_ref/model"
cp "$SRC/model/phet.py" "$HERE/_ref/model/phet.py"
: > "$HERE/_ref/model/__init__.py"
sha256sum "$HERE/_ref/model/phet.py" | cut -d' ' -f1 > "$HERE/_ref/SHA256"
echo "vendored $SRC/model/phet.py -> $HERE/_ref/model/phet.py ($(cat "$HERE/_ref/SHA256"))"

#!/usr/bin/env bash
# Installs the UNMODIFIED reference server under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box,
# which has no /root/reference) so that the wire-level tests and bench.py's `wire` leg can start it there:
#   baseline/_ref/{stt_server,stt_client,gen}   pip install --no-deps --target (the reference's own pyproject.toml)
#   baseline/_ref/_tree/{config,proto,tools}    the non-package files the server and its load generator read at run time
#                                               (config/server.yaml, config/model.yaml, proto/stt.proto, tools/bench/*.py)
# Nothing under baseline/_ref is product source and nothing of it is ever committed.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
DST="$REPO/baseline/_ref"
[ -d "$REF/stt_server" ] || { echo "no reference tree at $REF" >&2; exit 1; }
rm -rf "$DST" && mkdir -p "$DST"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/src"            # the build writes egg-info into the source tree; /root/reference is read-only
chmod -R u+w "$TMP/src"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" "$TMP/src" \
  || { echo "pip install failed; copying the three packages instead" >&2; cp -r "$TMP/src/stt_server" "$TMP/src/stt_client" "$TMP/src/gen" "$DST/"; }
mkdir -p "$DST/_tree/tools"
cp -r "$TMP/src/config" "$TMP/src/proto" "$DST/_tree/"
cp -r "$TMP/src/tools/bench" "$DST/_tree/tools/"
echo "reference installed under $DST"

#!/bin/bash
# Regenerates dsf2flac_b200/host/dsd_filter_bits.inc from the reference's filter definitions.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${DSF2FLAC_REFERENCE:-/root/reference}/src"
tmp="$(mktemp -d)"
g++ -O0 -I "$ref" -o "$tmp/gen" "$here/gen_filter_tables.cpp"
"$tmp/gen" > "$here/../dsf2flac_b200/host/dsd_filter_bits.inc"
rm -rf "$tmp"
echo "wrote dsf2flac_b200/host/dsd_filter_bits.inc"

#!/bin/bash
# Per-kernel census of the Blackwell-native (and legacy) instruction mnemonics in the shipped library:
#   UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UBLKCP = TMA (tensor / bulk copies), LDTM / STTM = tcgen05.ld / st,
#   HMMA = legacy mma.sync.  Usage: tools/sass_census.sh [out.txt]   (no GPU needed: cuobjdump reads the .so)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
LIB="$ROOT/misonet_b200/lib/libmisonet_b200.so"
OUT="${1:-$ROOT/profiles/r2_sass_census.txt}"
TMP="$(mktemp)"
cuobjdump -sass "$LIB" > "$TMP"
{
  echo "# SASS census of misonet_b200/lib/libmisonet_b200.so (git $(git -C "$ROOT" rev-parse --short HEAD 2>/dev/null), $(date -u +%Y-%m-%dT%H:%MZ))"
  echo "# cuobjdump -sass | per kernel: UTCHMMA (tcgen05.mma) UTMALDG (TMA tensor load) UBLKCP (TMA bulk copy) LDTM/STTM (tcgen05.ld/st) HMMA (mma.sync)"
  printf "%-64s %8s %8s %7s %6s %6s %6s\n" kernel UTCHMMA UTMALDG UBLKCP LDTM STTM HMMA
  awk '
    /Function :/ { if (name != "") print name, u, t, b, l, s, h; name=$3; u=t=b=l=s=h=0 }
    /UTCHMMA/ {u++} /UTMALDG/ {t++} /UBLKCP/ {b++} /LDTM/ {l++} /STTM/ {s++} /[^C]HMMA/ {h++}
    END { if (name != "") print name, u, t, b, l, s, h }' "$TMP" | while read -r name u t b l s h; do
      if [ "$u" != 0 ] || [ "$t" != 0 ] || [ "$b" != 0 ] || [ "$h" != 0 ]; then
        printf "%-64s %8s %8s %7s %6s %6s %6s\n" "$(echo "$name" | c++filt | sed -e 's/^void //' -e 's/miso:://g' -e 's/(anonymous namespace):://g' -e 's/(CUtensorMap.*//' -e 's/(.*//' | cut -c1-64)" "$u" "$t" "$b" "$l" "$s" "$h"
      fi
    done
  echo "# totals: UTCHMMA $(grep -c UTCHMMA "$TMP") UTMALDG $(grep -c UTMALDG "$TMP") UBLKCP $(grep -c UBLKCP "$TMP") LDTM $(grep -c LDTM "$TMP") STTM $(grep -c STTM "$TMP") HMMA (legacy mma.sync) $(grep -c "[^C]HMMA" "$TMP")"
} > "$OUT"
rm -f "$TMP"
cat "$OUT"

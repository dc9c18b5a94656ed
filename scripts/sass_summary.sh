#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (B200_PROFILING.md "What proves a
# Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA tensor copies -> UTMALDG, TMA bulk
# copies -> UBLKCP, cp.async -> LDGSTS, mma.sync -> HMMA/IMMA.  Usage: scripts/sass_summary.sh > profiles/rNN_sass_summary.txt
LIB=${1:-unnamed-rust-sdr_b200/lib/libsdr_b200.so}
echo "# SASS mnemonic counts per kernel of $LIB ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -1))"
echo "# columns: UTCIMMA UTCHMMA LDTM UTMALDG UBLKCP LDGSTS HMMA IMMA SYNCS  kernel"
cuobjdump -sass "$LIB" 2>/dev/null | awk '
function flush() { if (name != "") printf "%7d %7d %5d %7d %6d %6d %5d %5d %5d  %s\n", c["UTCIMMA"], c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UBLKCP"], c["LDGSTS"], c["HMMA"], c["IMMA"], c["SYNCS"], name }
/Function :/ { flush(); name = $3; delete c; next }
{ for (k in K) if (index($0, k)) { if ((k == "HMMA" || k == "IMMA") && index($0, "UTC")) continue; c[k]++ } }
BEGIN { split("UTCIMMA UTCHMMA LDTM UTMALDG UBLKCP LDGSTS HMMA IMMA SYNCS", a, " "); for (i in a) K[a[i]] = 1 }
END { flush() }' | while read -r l; do set -- $l; n=$(echo "${10}" | c++filt 2>/dev/null | sed -e 's/sdr::(anonymous namespace):://g' -e 's/((anonymous namespace)::[A-Za-z]*)//g' | cut -c1-90); s=$(( $1 + $2 + $3 + $4 + $5 + $6 + $7 + $8 )); [ "$s" -gt 0 ] && printf "%7d %7d %5d %7d %6d %6d %5d %5d %5d  %s\n" $1 $2 $3 $4 $5 $6 $7 $8 $9 "$n"; done

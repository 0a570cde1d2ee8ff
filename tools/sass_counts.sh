#!/bin/bash
# Opcode evidence that the sketch engines are Blackwell-native (tcgen05 / TMEM / TMA): counts of the SASS mnemonics in
# libgpca.so.  usage: tools/sass_counts.sh > profiles/rN_sass_opcodes.txt
SO="$(dirname "$0")/../genomic_pca_b200/libgpca.so"
echo "# cuobjdump -sass $(basename "$SO")  ($(date -u +%F), nvcc $(nvcc --version | grep -o 'V[0-9.]*'))"
echo "# tcgen05.mma -> UTC*MMA ; tcgen05.ld/st -> LDTM/STTM ; TMA -> UTMALDG / UBLKCP ; tcgen05.commit -> UTCBAR ; mbarrier -> SYNCS"
cuobjdump -sass "$SO" | grep -oE '\b(UTCIMMA|UTCHMMA|UTCQMMA|UTCOMMA|STTM|LDTM|UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|UTCBAR[.A-Z0-9]*|SYNCS[.A-Z0-9]*|HMMA[.A-Z0-9]*|IMMA[.A-Z0-9]*|POPC|LOP3[.A-Z0-9]*|REDUX[.A-Z0-9]*|SHFL[.A-Z0-9]*)\b' | sort | uniq -c | sort -k1,1nr
echo "# per kernel (functions that contain tensor-core or TMA instructions)"
cuobjdump -sass "$SO" | awk '/Function :/ {fn=$3} /UTC[A-Z]*MMA|STTM|LDTM|UTMALDG|UBLKCP/ {c[fn]++} END {for (f in c) print c[f], f}' | sort -k1,1nr | c++filt | cut -c1-160

#!/bin/bash
# SASS opcode summary of the in-tree library (what proves the Blackwell-native path: LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,
# SYNCS = mbarrier, DFMA/DADD/DMUL = the fp64 pipe).  Runs without a GPU:  bash tools/sass_summary.sh > profiles/<tag>_sass_summary.txt
LIB=${1:-cosmoprimo_b200/libcpfftlog.so}
echo "# cuobjdump -sass $LIB | opcode counts ($(date -u +%F))"
cuobjdump -sass $LIB > /tmp/cpf_sass.txt
echo "## whole library"
for op in LDTM STTM UBLKCP UTMALDG SYNCS DFMA DADD DMUL DSETP LDS STS LDG STG ATOM RED BAR SHFL HMMA UTC; do
  printf "%-8s %d\n" $op $(grep -c "^\s*/\*[0-9a-f]*\*/\s*\(@!\?U\?P[0-9T] \)\?$op" /tmp/cpf_sass.txt)
done
echo "## per kernel (DFMA / DADD / DMUL / LDS+STS / LDTM / UBLKCP)"
awk '/Function : /{name=$3} /DFMA/{f[name]++} /DADD/{a[name]++} /DMUL/{m[name]++} /[^A-Z](LDS|STS)/{s[name]++} /LDTM/{t[name]++} /UBLKCP/{u[name]++} END{for(n in f) printf "%-110s %6d %6d %6d %6d %5d %4d\n", substr(n,1,110), f[n], a[n], m[n], s[n], t[n], u[n]}' /tmp/cpf_sass.txt | sort | c++filt 2>/dev/null

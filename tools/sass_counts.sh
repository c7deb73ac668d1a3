#!/bin/bash
# SASS evidence (no GPU needed): per kernel of libszn.so, how many tcgen05 / TMEM / TMA instructions it contains.
#   bash tools/sass_counts.sh > profiles/r02_sass_counts.txt
SO=zeroshotsemanticsegmentation_b200/libszn.so
echo "# cuobjdump -sass $SO ($(sha256sum $SO | cut -c1-16)), $(date -u +%Y-%m-%dT%H:%MZ)"
echo "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add, UTCBAR = tcgen05.commit,"
echo "# ELECT = elect.sync, BRA.U.ANY = the per-thread serialisation loop ptxas wraps around uniform-datapath instructions issued under"
echo "# a divergent branch (0 everywhere since round 2: every UTCHMMA / UTMALDG is issued by an elected lane of a converged warp)"
cuobjdump -sass $SO | awk '
  /Function :/ { name=$3; order[++n]=name }
  /UTCHMMA/ {c[name,"UTCHMMA"]++} /LDTM/ {c[name,"LDTM"]++} /UTMALDG/ {c[name,"UTMALDG"]++} /UTMASTG/ {c[name,"UTMASTG"]++}
  /UTMAREDG/ {c[name,"UTMAREDG"]++} /UTCBAR/ {c[name,"UTCBAR"]++} /ELECT/ {c[name,"ELECT"]++} /BRA.U.ANY/ {c[name,"BRAUANY"]++}
  /UTCATOMSWS|UTCALLOC/ {c[name,"TMEMALLOC"]++}
  END { printf "%-8s %-6s %-8s %-8s %-9s %-7s %-6s %-10s  %s\n","UTCHMMA","LDTM","UTMALDG","UTMASTG","UTMAREDG","UTCBAR","ELECT","BRA.U.ANY","kernel";
        for (i=1;i<=n;i++){k=order[i]; t=c[k,"UTCHMMA"]+c[k,"LDTM"]+c[k,"UTMALDG"]+c[k,"UTMASTG"]+c[k,"UTMAREDG"];
          if (t>0) printf "%-8d %-6d %-8d %-8d %-9d %-7d %-6d %-10d  %s\n",c[k,"UTCHMMA"],c[k,"LDTM"],c[k,"UTMALDG"],c[k,"UTMASTG"],c[k,"UTMAREDG"],c[k,"UTCBAR"],c[k,"ELECT"],c[k,"BRAUANY"],k} }' | c++filt | sed 's/(CUtensorMap_st.*//'

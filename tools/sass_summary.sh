#!/bin/bash
# SASS evidence for profiles/: per hot kernel the opcode histogram and the instructions that prove TMA / PDL / mbarrier /
# cp.async use.  tools/sass_summary.sh <object> <mangled-name substring> <out>
OBJ=$1; PAT=$2; OUT=$3
cuobjdump -sass "$OBJ" 2>/dev/null | awk '/Function : /{name=$3} {print name"\t"$0}' | grep "$PAT" > /tmp/sass_one.txt
name=$(head -1 /tmp/sass_one.txt | cut -f1)
{
  echo "# cuobjdump -sass $(basename $OBJ), function $name"
  echo "# $(cu++filt $name 2>/dev/null | head -1)"
  echo "# instructions: $(grep -cE '/\*[0-9a-f]{4}\*/' /tmp/sass_one.txt)"
  echo "# registers / stack / shared (cuobjdump -res-usage):"
  cuobjdump -res-usage "$OBJ" 2>/dev/null | grep -A1 "$name" | tail -1 | sed 's/^/#   /'
  echo "## opcode histogram"
  cut -f2 /tmp/sass_one.txt | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*//; s/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn
  echo "## TMA / mbarrier / programmatic-launch / async-copy instructions"
  cut -f2 /tmp/sass_one.txt | grep -E "UTMALDG|UTMASTG|UTMACMDFLUSH|SYNCS|ACQBULK|LDGSTS|LDGDEPBAR|UTMAPF|CCTL|ERRBAR|MEMBAR|FENCE" | sed -E 's/^\s+//; s/\s*\/\*[0-9a-f]{16}\*\///' 
} > "$OUT"
wc -l "$OUT"

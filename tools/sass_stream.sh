#!/bin/bash
# usage: tools/sass_stream.sh <lib.so> <kernel-name-substring>   -- compressed opcode stream of one kernel (Fn = n consecutive FFMA2)
cuobjdump -sass "$1" 2>/dev/null | awk -v k="$2" '/Function : /{f=index($0,k)>0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{for(i=2;i<=NF;i++){if($i ~ /^[A-Z][A-Z0-9_.]+$/){print $i; break}}}' | sed 's/\..*//' | awk '{ if ($1=="FFMA2") c++; else { if (c) printf "F%d ", c; c=0; printf "%s ", $1 } } END {print ""}' | fold -w 200

#!/bin/sh
# weaken.sh IN.o OUT.o "SYM ..." "SYM=ALIAS ..."   -- TEST INFRASTRUCTURE (link-time drop-in demonstration).
# Copies an UNMODIFIED reference object, making the listed global symbols weak (the glue objects carry the strong
# definitions; the reference objects are -fPIC, so every call goes through the symbol and binds to the strong one) and
# adding, for every SYM=ALIAS pair, a second global name ALIAS at SYM's address: the reference's own body stays
# reachable for a glue function that does some work of its own and then runs it.
set -e
in=$1; out=$2; weak=$3; alias=$4
args=""
for s in $weak; do args="$args --weaken-symbol=$s"; done
for pair in $alias; do
    sym=${pair%%=*}; name=${pair#*=}
    line=$(objdump -t "$in" | awk -v s="$sym" '$NF==s && $2=="g" {print $1, $4}')
    off=${line%% *}; sec=${line#* }
    if [ -z "$off" ] || [ -z "$sec" ]; then echo "weaken.sh: $sym not found in $in" >&2; exit 1; fi
    args="$args --add-symbol $name=$sec:0x$off,function,global"
done
objcopy $args "$in" "$out"

#!/bin/bash
# usage: run_variant.sh <variant.so> <command...>  : temporarily swaps the library (GPU box scratch copy only)
cp agdiff_b200/libagdiff_b200.so /tmp/lib_backup.so
cp "$1" agdiff_b200/libagdiff_b200.so
shift
"$@"
cp /tmp/lib_backup.so agdiff_b200/libagdiff_b200.so

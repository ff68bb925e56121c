#!/bin/bash
# Thread-count sweep of the three join algorithms, in the format of the reference's
# scripts/tput-scalability.sh: every run appends "<algo> <nthreads> <rep> " followed by what the
# binary writes to stderr (five per-stage times, NUM-TUPLES, TOTAL-TIME-USECS, TUPLES-PER-SECOND,
# src/joins/joincommon.c:176-196) to $RESULTS/$OUTPUTFILE.txt, and stdout to the .log file.
# The reference's script hard-codes its binary, the workload and the thread list; here they come
# from the environment so that the same sweep runs against either build and at test sizes:
#
#   BIN=oracle/_ref/sortmergejoins_dropin NTUPLES=128000000 scripts/tput-scalability.sh   # reference main.c on the GPU library
#   BIN=avx-sort-merge-joins_b200/sortmergejoins scripts/tput-scalability.sh              # this repo's driver
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
BIN="${BIN:-$HERE/../avx-sort-merge-joins_b200/sortmergejoins}"
RESULTS="${RESULTS:-$HERE/../measurements}"
NTUPLES="${NTUPLES:-128000000}"
OUTPUTFILE="${OUTPUTFILE:-tput-WorkloadB}"
THREADS="${THREADS:-64 32 16 8 4 2 1}"
ALGOS="${ALGOS:-m-pass m-way mpsm}"
REPS="${REPS:-3}"
mkdir -p "$RESULTS"
OUTFILE="$RESULTS/$OUTPUTFILE.txt"
LOGFILE="$RESULTS/$OUTPUTFILE.log"

sweep() {  # $1 = label suffix, remaining args = extra options
    local suffix="$1"; shift
    echo "#ALGO NTHREADS RUNNO PARTCYC SORTCYC MERGE1CYC MERGERESTCYC MJOINCYC NUMTUP USECS TPUT" >> "$OUTFILE"
    for ALGO in $ALGOS; do
        for NTHREADS in $THREADS; do
            for REP in $(seq 1 "$REPS"); do
                EXE="$BIN --algo=$ALGO --nthreads=$NTHREADS --r-size=$NTUPLES --s-size=$NTUPLES $*"
                echo "========================================" >> "$LOGFILE"
                echo "$EXE" >> "$LOGFILE"
                echo -n "$ALGO$suffix $NTHREADS $REP " >> "$OUTFILE"
                $EXE >> "$LOGFILE" 2>> "$OUTFILE"
                echo "" >> "$OUTFILE"
            done
        done
    done
}

echo "========== Algorithms implemented with AVX ==========" >> "$OUTFILE"
sweep ""
echo "========== Scalar algorithms without AVX ==========" >> "$OUTFILE"
sweep "-scalar" --scalarmerge --scalarsort

cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for t in memcheck synccheck racecheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_case.py > gpurun_out/san_$t.log 2>&1; echo "$t rc=$?"; grep "SUMMARY\|sanitize case done" gpurun_out/san_$t.log | tail -3; done
grep "Race reported" gpurun_out/san_racecheck.log | sed 's/+0x[0-9a-f]*//g; s/ and .*//' | cut -c1-230 | sort | uniq -c | sort -rn | head -30

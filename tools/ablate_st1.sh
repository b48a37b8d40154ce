for f in 0 8192 32768 40960 16384; do echo "== kb32 flags $f"; timeout 40 python tools/trace_tma.py st1 $f 2>&1 | grep -E "us \(|MMA   issue|MMA   tile period|EPI   items"; done
echo "== kb32 mt1"; timeout 40 python tools/trace_tma.py st1 0 1 2>&1 | grep -E "us \(|MMA   issue|MMA   tile period|EPI   items"
export FTC_NO_KB32=1
for f in 0 8192 32768; do echo "== sw128 flags $f"; timeout 40 python tools/trace_tma.py st1 $f 2>&1 | grep -E "us \(|MMA   issue|MMA   tile period|EPI   items"; done

# Sweep of the scan's tuning knobs on configs[1] (GPU box): python tools/run_once.py scan 30 <option=value ...>
for opts in "" "scan_tile_bytes=36864" "scan_tile_bytes=40960 scan_stages=2" "scan_tile_bytes=49152 scan_stages=2" "scan_stage_buf_bytes=8192" "scan_stage_buf_bytes=2048" "scan_chunk_tiles=8" "scan_pdl=0"; do
  echo "== $opts"; python tools/run_once.py scan 30 $opts 2>&1 | tail -1 | python -c "
import sys,re
l=sys.stdin.read()
ms=[float(x) for x in re.findall(r\"'([0-9.]+)'\", l)]
ms=sorted(ms[5:])
print('novel', re.search(r'novel (\d+)', l).group(1), 'median %.4f min %.4f mean %.4f' % (ms[len(ms)//2], ms[0], sum(ms)/len(ms)))
"
done

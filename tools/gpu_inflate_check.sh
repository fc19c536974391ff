#!/bin/bash
# GPU box: the host reader with and without the GPU inflate engine on one synthetic BAM (bamstat = reader only)
python - <<'PY'
import sys,subprocess,json,time,os
sys.path.insert(0,'.')
from synth import synth as S
scale=float(os.environ.get("SCALE","0.15"))
w=S.make_workload(3,scale=scale)
t=time.time(); S.write_bam(w,'/tmp/g.bam',with_seq=True); print('write',round(time.time()-t,1),'s',os.path.getsize('/tmp/g.bam')/1e9,'GB')
cli='inquistr_b200/bin/inquistr-b200'
for args in ([], ['--gpu','0'], [], ['--gpu','0']):
    t=time.time()
    r=subprocess.run([cli,'bamstat','/tmp/g.bam']+args,capture_output=True,text=True)
    d=json.loads(r.stdout); print(args,'wall',round(time.time()-t,2),{k:d[k] for k in ('records','seconds','inflate_GBps','blocks_fast','blocks_gpu','bytes_gpu')}, r.stderr[-200:])
PY

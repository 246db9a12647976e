import json, subprocess, time, sys
V = json.load(open('/root/repo/tests/golden/ckzg_vectors.json'))
h = lambda s: bytes.fromhex(s[2:])
recs=[]; exp=[]
for c in V['verify_kzg_proof']:
    a=[h(c[k]) for k in ('commitment','z','y','proof')]
    if [len(x) for x in a]!=[48,32,32,48]: continue
    recs.append(b"".join(a)); exp.append({True:'1',False:'0',None:'2'}[c['output']])
t=time.time()
out = subprocess.run([sys.argv[1],'/root/repo/kzg_rs_b200/data/mainnet_setup.bin'], input=b"".join(recs), capture_output=True).stdout.decode().strip()
print(len(recs), 'time', time.time()-t)
print(out); print("".join(exp)); print('MATCH' if out=="".join(exp) else 'MISMATCH')

import csv,sys
for f in sys.argv[1:]:
    rows=[r for r in csv.reader(open(f)) if len(r)>5]
    hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
    data=[(r[ik],float(r[iv].replace(',',''))) for r in rows[1:]]
    last=data[-9:]
    print(f, ' '.join('%s=%.0f'%(d[0].split('::')[-1][:10],d[1]/1000) for d in last), 'total %.0f'%(sum(d[1] for d in last)/1000))

import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]
ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki].split('(')[0][:28]),{})[r[mi]]=r[vi]
for k,v in d.items():
    t=float(v.get('gpu__time_duration.sum','0').replace(',',''))/1e3
    rd=float(v.get('dram__bytes_read.sum','0').replace(',',''))/1e6
    wr=float(v.get('dram__bytes_write.sum','0').replace(',',''))/1e6
    print(f"{k[0]:>3} {k[1]:<28} {t:9.1f} us  dram rd {rd:8.1f} MB  wr {wr:8.1f} MB  " + " ".join(f"{a.split('__')[-1][:18]}={b}" for a,b in v.items() if 'lts' in a or 'l1tex' in a or 'smsp' in a))

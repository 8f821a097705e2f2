import re, subprocess, sys
log=open(sys.argv[1]).read()
ents=re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*\n.*Used (\d+) registers", log)
for name,stack,spill,regs in ents:
    dem=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()
    dem=re.sub(r'fvdbm::','',dem); dem=re.sub(r'\(.*','',dem).replace('void ','')
    if len(sys.argv)<3 or re.search(sys.argv[2],dem):
        print(f"{dem:55s} regs={regs:>3s} stack={stack} spill={spill}")

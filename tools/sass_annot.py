"""Annotate the SASS of one kernel with CUDA source lines (needs -lineinfo): python tools/sass_annot.py <obj|so> <kernel substring> > out.txt
Each output line: source line <TAB> address <TAB> instruction.  Used to count instructions on the hot path without a GPU."""
import os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, stdout=subprocess.DEVNULL, check=True)
for cb in sorted(os.listdir(td)):
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(td, cb)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    secs = re.split(r"\n(?=\.text\.)", txt)
    for sec in secs:
        head = sec.split("\n", 1)[0]
        if not head.startswith(".text.") or pat not in head:
            continue
        cur = None
        for l in sec.split("\n"):
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = int(m.group(2)); continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", l)
            if m:
                print("%s\t%s\t%s" % (cur, m.group(1), m.group(2)))
            elif re.match(r"\s*\.L_x_\d+:", l):
                print("\t\t" + l.strip())
        sys.exit(0)
sys.exit("kernel not found")

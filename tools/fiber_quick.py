import sys, os, json, hashlib, subprocess, tempfile
ROOT = os.getcwd()
sys.path.insert(0, ROOT + "/tests/golden"); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT)
import cases as gc
gold = json.load(open(ROOT + "/tests/golden/paf_md5.json"))
with tempfile.TemporaryDirectory() as td:
    cs = dict(gc.build_cases(td))
    for name, env in (("syn_ont", {}), ("sr_paired", {}), ("tandem_iter64", {}), ("syn_ccs", {"MM2B_FIBER_ASYNC": "1"}), ("sr_paired", {"MM2B_FIBER_ASYNC": "1"})):
        p = subprocess.run([ROOT + "/oracle/_ref/minimap2-fiber-b200", "-t", "256"] + cs[name], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **env), timeout=30)
        print(name, env, p.returncode, hashlib.md5(p.stdout).hexdigest() == gold[name]["md5"], flush=True)

"""Compare two builds of libatspeed_b200.so kernel by kernel (instruction text + encodings from `cuobjdump -sass`,
whitespace-normalised): which kernels are new, gone, or changed.  Used when adding opt-in experimental kernels to show
that the tested default kernels are untouched.   usage: python tools/sass_diff.py old.so new.so [old_name=new_name ...]"""
import re
import subprocess
import sys


def kernels(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        mangled, body = f.split("\n", 1)
        dem = subprocess.run(["cu++filt", mangled.strip()], capture_output=True, text=True).stdout.strip() or mangled.strip()
        dem = re.sub(r"\((int|bool)\)", "", dem).replace("void atspeed::", "").replace("atspeed::", "")
        dem = re.split(r"\((?![^<]*>)", dem)[0]
        out[dem] = [re.sub(r"\s+", " ", l).strip() for l in body.splitlines() if "/*" in l]
    return out


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    rename = dict(a.split("=") for a in sys.argv[3:])
    same = changed = 0
    for k, v in sorted(old.items()):
        k2 = rename.get(k, k)
        if k2 not in new:
            print("GONE    ", k)
        elif new[k2] != v:
            print("CHANGED ", k, "->", k2)
            changed += 1
        else:
            same += 1
    for k in sorted(set(new) - {rename.get(k, k) for k in old}):
        print("NEW     ", k)
    print(f"{same} kernels identical, {changed} changed")


if __name__ == "__main__":
    main()

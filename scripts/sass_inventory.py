"""Which Blackwell-native instructions each kernel of libnanodec.so contains (cuobjdump -sass; works without a GPU).
usage: python scripts/sass_inventory.py [--md]   (--md prints the table kept in profiles/r01_sass_inventory.md)"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nanodecoder_b200", "libnanodec.so")
WANT = ["UTC*MMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "STAS", "SYNCS", "FFMA2", "HMMA"]


def inventory(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = out.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if not m:
            continue
        op = m.group(1)
        key = "UTC*MMA" if re.fullmatch(r"UTC[A-Z]*MMA", op) else op
        if key in WANT:
            cur[key] += 1
    return out


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in r:
        n = re.sub(r"(nd::)?\(anonymous namespace\)::", "", n)
        n = re.sub(r"^void ", "", n)
        n = re.sub(r"\(.*$", "", n)                       # argument list (the namespace parenthesis is gone already)
        short.append(n.replace("nd::", ""))
    return short


if __name__ == "__main__":
    inv = inventory()
    names = demangle(list(inv))
    rows = [(n, c) for n, c in zip(names, inv.values()) if any(c[k] for k in WANT if k not in ("SYNCS",))]
    if "--md" in sys.argv:
        print("| kernel | " + " | ".join("`%s`" % k for k in WANT) + " |\n|---|" + "---:|" * len(WANT))
        for n, c in sorted(rows):
            print("| `%s` | " % n + " | ".join(str(c[k]) if c[k] else "" for k in WANT) + " |")
    else:
        for n, c in sorted(rows):
            print("%-70s %s" % (n[:70], dict(c)))

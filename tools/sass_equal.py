"""Are the kernels of two builds of libvfk.so the same machine code?

    python tools/sass_equal.py old/libvfk.so vafator_b200/libvfk.so [--ignore REGEX]

Compares `cuobjdump -sass` kernel by kernel (addresses and encodings stripped; trailing `Lb0` (false) template arguments
of the pileup kernels are dropped from the names, so that a build that only ADDS opt-in instantiations -- some `Lb1` --
compares equal to the build before it).  Used to show that the staged, not-yet-verified kernel variants leave the
GPU-verified default path untouched."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = re.sub(r"(pileup_(?:warp|column)_kernelILi\d+)(?:ELb0)+E", r"\1E", m.group(1))
            body[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body[name].append(re.sub(r"/\*[0-9a-f]{4}\*/|/\* 0x[0-9a-f]+ \*/", "", line).strip())
    return body


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ignore = re.compile(sys.argv[sys.argv.index("--ignore") + 1]) if "--ignore" in sys.argv else re.compile(r"Lb1")
    if "--ignore" in sys.argv:
        args.remove(sys.argv[sys.argv.index("--ignore") + 1])
    a, b = kernels(args[0]), kernels(args[1])
    a = {k: v for k, v in a.items() if not ignore.search(k)}
    b = {k: v for k, v in b.items() if not ignore.search(k)}
    bad = sorted(k for k in set(a) | set(b) if a.get(k) != b.get(k))
    print("%d kernels compared, %d instructions, %d differ" % (len(set(a) & set(b)), sum(map(len, a.values())), len(bad)))
    for k in bad:
        print("  differs:", k, len(a.get(k, [])), "->", len(b.get(k, [])))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()

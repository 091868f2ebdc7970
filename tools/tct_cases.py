import sys, tempfile
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import tc_probe as t
tmp = tempfile.mkdtemp()
for (T, D) in [(256, 16), (64, 16), (128, 16), (272, 16), (64, 8), (256, 8), (512, 8)]:
    try:
        t.one(T, D, 300032, 2, tmp)
    except Exception as e:
        print(T, D, "ERR", e)
t.one(256, 16, 8192 + 256, 1, tmp)
t.one(256, 16, 40000, 3, tmp)
t.one(256, 16, 1 << 20, 2, tmp, full_range=True)
t.one(256, 8, 1 << 20, 2, tmp, full_range=True)

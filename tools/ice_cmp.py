import sys, numpy as np
a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
bad = [k for k in a.files if not np.array_equal(a[k], b[k])]
print("bitwise equal" if not bad else f"DIFFER: {bad}", sys.argv[1], sys.argv[2])

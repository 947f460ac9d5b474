import numpy as np
a = np.load("/tmp/res_off.npz"); b = np.load("/tmp/res_on.npz")
print("ids equal:", np.array_equal(a["ids"], b["ids"]), "dist equal:", np.array_equal(a["d"].view(np.uint32), b["d"].view(np.uint32)))
if not np.array_equal(a["ids"], b["ids"]):
    bad = np.nonzero((a["ids"] != b["ids"]).any(axis=1))[0]
    print("bad rows", len(bad), bad[:10])

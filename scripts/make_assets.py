"""Convert the reference's only hot-path data fixture into the packaged asset.

Reads /root/reference/assets/contact_zones.pkl (loaded by the reference at
handobjectdatasets/contactutils.py:12-14,45; consumed at
mano_train/networks/branches/contactloss.py:262-274) and writes
obman_train_b200/assets/contact_zones.npz with the same content in a pickle-free layout:

  verts  (778,3) float64   MANO template vertices, metres, zero-mean
  faces  (1538,3) int64    MANO template faces
  zone_ids   (115,) int64  concatenated per-zone hand-vertex ids
  zone_ptr   (7,)  int64   CSR offsets into zone_ids (6 zones: 19/28/19/7/25/17 ids)

Run in the build container only (the GPU box has no /root/reference).
"""
import os
import pickle

import numpy as np

SRC = "/root/reference/assets/contact_zones.pkl"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "obman_train_b200", "assets", "contact_zones.npz")


def main():
    with open(SRC, "rb") as f:
        data = pickle.load(f, encoding="latin1")
    zones = data["contact_zones"]
    ids, ptr = [], [0]
    for k in sorted(zones.keys()):
        ids.extend(int(i) for i in zones[k])
        ptr.append(len(ids))
    np.savez_compressed(
        DST,
        verts=np.asarray(data["verts"], dtype=np.float64),
        faces=np.asarray(data["faces"]).astype(np.int64),
        zone_ids=np.asarray(ids, dtype=np.int64),
        zone_ptr=np.asarray(ptr, dtype=np.int64),
    )
    print("wrote", DST, ptr)


if __name__ == "__main__":
    main()

"""Stage-by-stage GPU vs oracle comparison on one fuzz stream (debug helper)."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import fuzz_common as fz
from oracle import c_oracle
from p265_b200 import deblock_api, intra_host, packer, sao_api, scaling_list
from p265_b200.engine import Engine
from p265_b200.picture import PicGeom
c_oracle.build()
name = sys.argv[1]
cfg = fz.manifest()[name]
imgs, sps, pps = fz.parse(name, cfg)
eng = Engine(0)
for p, img in enumerate(imgs):
    w, h, bd, cl = int(sps.pic_width_in_luma_samples), int(sps.pic_height_in_luma_samples), int(sps.bit_depth_y), int(sps.ctb_log2_size_y)
    batch = packer.pack_pictures([img], sps, scaling_list.active_table(sps, pps))
    r_g, r_o = eng.residual(batch), c_oracle.residual_batch(batch)
    print("pic", p, "residual equal", np.array_equal(r_g, r_o))
    rec = intra_host.reconstruct_intra_picture(img, sps, pps, [batch.geom.plane_view(r_o, 0, c) for c in range(3)])
    geom = PicGeom(w, h, 1, bd, bd)
    buf = np.zeros(geom.total_elems(), np.uint8 if bd <= 8 else np.uint16)
    for c in range(3): geom.plane_view(buf, 0, c)[:] = rec[c]
    blk, ctb = deblock_api.edge_map_from_picture(img, sps, pps)
    d_g, d_o = eng.deblock(buf, geom, cl, blk, ctb), c_oracle.deblock_batch(buf, geom, cl, blk, ctb)
    for c in range(3):
        a, b = geom.plane_view(d_g, 0, c), geom.plane_view(d_o, 0, c)
        bad = np.argwhere(a != b)
        print("  deblock comp", c, "mismatches", len(bad), bad[:6].tolist(), [(int(a[y, x]), int(b[y, x]), int(geom.plane_view(buf, 0, c)[y, x])) for y, x in bad[:6]])
    nf = sao_api.no_filter_from_picture(img, sps)
    for label, avail, nfl in (("spec", sao_api.availability_from_picture(img, sps, pps), nf), ("lav", fz.lav_sao_avail(img, sps), nf), ("nonf", fz.lav_sao_avail(img, sps), None)):
        params = packer.sao_params_from_picture(img, sps, avail)
        s_g, s_o = eng.sao(d_o, geom, cl, params, no_filter=nfl), c_oracle.sao_batch(d_o, geom, cl, params, nfl)
        for c in range(3):
            a, b = geom.plane_view(s_g, 0, c), geom.plane_view(s_o, 0, c)
            bad = np.argwhere(a != b)
            if len(bad):
                print("  sao", label, "comp", c, "mismatches", len(bad), bad[:6].tolist(), [(int(a[y, x]), int(b[y, x])) for y, x in bad[:6]])

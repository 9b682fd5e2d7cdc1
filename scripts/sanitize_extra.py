"""Second leg of scripts/sanitize.sh: a small many-light scene through every kernel family (fused and split G-buffer / phase A,
spatial passes, PTDirect, tone-map, the reference-order walk and strips with halo copies), so that compute-sanitizer sees them."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run():
    import restir_b200 as rb
    from restir_b200 import scenes

    rb.init(0)
    sd = scenes.procedural(2, 6000, 300, (160, 96))
    sc = rb.Scene.from_arrays(sd)
    base = rb.Camera.from_scene(sd)
    for exact in (False, True):
        sc.set_traversal(exact)
        for fuse in (True, False):
            fr = sc.frame(160, 96)
            fr.set_fusion(fuse)
            prm = rb.default_params(reuse=3, radius=12.0, passes=2)
            for k in range(3):
                cam = base.orbit(k)
                fr.gbuffer_render(cam); fr.restir_direct(cam, prm, k, 0); fr.gbuffer_update(cam)
            fr.pathtrace_direct(base, 7, 0)
            fr.tonemap(rb.TONEMAP_ACES, 1.0)
            fr.read("ldr"); fr.read("reservoir"); fr.read("light_index"); fr.read("normal")
            fr.close()
    sc.set_traversal(False)
    a, b = sc.frame(160, 96, rows=(0, 48), halo=16), sc.frame(160, 96, rows=(48, 96), halo=16)
    prm = rb.default_params(reuse=3, radius=12.0)
    for k in range(3):
        cam = base.orbit(k)
        for s in (a, b):
            s.set_halo_render(False)
            s.gbuffer_render(cam); s.restir_phase_a(cam, prm, k, 0)
        for plane in ("geom_cur", "matid_cur", "resv_temp", "resv_out"):
            b.copy_rows_from(a, plane, 32, 48)
            a.copy_rows_from(b, plane, 48, 64)
        for s in (a, b):
            s.restir_phase_b(cam, prm, k, 0); s.gbuffer_update(cam)
    assert a.halo_miss() == 0 and b.halo_miss() == 0
    a.close(); b.close(); sc.close()
    print("sanitize_extra ok")


if __name__ == "__main__":
    run()

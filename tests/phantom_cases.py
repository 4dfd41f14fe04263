"""Phantom generator cases shared by the CPU (oracle pinning) and GPU (parity) tests.  Keys are oracle.pyphantom keyword arguments."""
CYLINDER, SPHERE, TWOPOOLS = 0, 1, 2

CASES = {
    # the demo notebooks' cylinder recipe (r = 8 µm, BVF 4 %, Y = 0.78, perpendicular to B0), small grid
    "cyl_bold": dict(shape=CYLINDER, fov_um=100.0, resolution=64, radius_um=8.0, volume_fraction=4.0, Y=0.78, orientation_deg=90.0, seed=0),
    # random radii, oblique B0, odd resolution (exercises the broadcast tail: 50^3 is not a multiple of 4096)
    "cyl_random_oblique": dict(shape=CYLINDER, fov_um=100.0, resolution=50, radius_um=-12.0, volume_fraction=10.0, Y=0.6, orientation_deg=35.0, seed=3),
    # parallel to B0: projected B0 is the zero vector (phantom_cylinder.cpp:200-201), sin^2 = 0
    "cyl_parallel": dict(shape=CYLINDER, fov_um=200.0, resolution=96, radius_um=5.0, volume_fraction=6.0, Y=0.85, orientation_deg=0.0, seed=7),
    # mask only (-y -1)
    "cyl_mask_only": dict(shape=CYLINDER, fov_um=80.0, resolution=40, radius_um=6.0, volume_fraction=5.0, Y=-1.0, seed=2),
    # anisotropic-looking numbers: fov not a multiple of the resolution, many small cylinders -> several smem batches (> 256 shapes)
    "cyl_many": dict(shape=CYLINDER, fov_um=333.0, resolution=72, radius_um=-6.0, volume_fraction=12.0, Y=0.7, orientation_deg=60.0, seed=11),
    # the dwi notebook's sphere recipe (random radii, 30-40 %), small grid; > 256 spheres
    "sph_random": dict(shape=SPHERE, fov_um=100.0, resolution=64, radius_um=-10.0, volume_fraction=30.0, Y=0.78, seed=0),
    "sph_fixed": dict(shape=SPHERE, fov_um=60.0, resolution=48, radius_um=4.0, volume_fraction=10.0, Y=0.5, seed=5),
    "sph_mask_only": dict(shape=SPHERE, fov_um=60.0, resolution=45, radius_um=-6.0, volume_fraction=20.0, Y=-1.0, seed=9),
    "twopools_odd": dict(shape=TWOPOOLS, fov_um=10.0, resolution=15),
    "twopools_even": dict(shape=TWOPOOLS, fov_um=10.0, resolution=32),
}

"""Copies the reference's known-answer DATA (literature / analytic curves its tests compare against) into fixtures.

/root/reference does not exist on the GPU box, so the data sets the reference's own post-processing compares
against travel as small .npz files.  They are data, not code:

  prosperetti_capillary.npz   test/small_test/multiphase/capillary_wave/prosperetti.csv -- Prosperetti's analytic
                              solution of the viscous capillary wave (omega_0 t, maximum interface amplitude); the
                              curve capillary_wave/postpro.py:92-101 measures its result against.
  rising_bubble_com_ref.npz   test/small_test/multiphase/rising_bubble/com_ref.txt -- the benchmark solution of the
                              rising-bubble test case 1 (Hysing et al.): time, centre-of-mass height and rise velocity
                              (columns 0, 3, 4), the curves rising_bubble/postpro.py:55-75 plots its result against; t2, yc2, uc2: the same
                              columns of com_ref_2.txt, test case 2 (density ratio 1000, sigma = 1.96; postpro_2.py).

  ghia_cavity_re1000.npz      test/small_test/navier_stokes/lid_driven/uref, vref -- Ghia, Ghia & Shin's centreline
                              velocities of the lid-driven cavity at Re = 1000 (17 points each, coordinate - 0.5), the
                              points lid_driven/postpro.py:57-78 plots its profiles against.

  (lid3d_re1000_64.npz, written by make_lid3d.py, carries test/large_test/lid3D/Uref.csv and Vref.csv -- Ku et al.'s
   centreline velocities of the cubic cavity at Re = 1000 -- next to the oracle run they are compared with.)

  shear_drop_basilisk.npz     test/small_test/multiphase/shear_drop/reference/Re1Ca02b.csv, Re1Ca04b.csv, Re1Ca09b.csv --
                              the Basilisk deformation curves D(t) of a drop in shear flow at Re = 1, Ca = 0.2, 0.4,
                              0.9 (columns t, D), the points shear_drop/postpro.py:44-52 plots its curves against, and the drop
                              contours shapeRe1Ca0*b.txt it draws its final vof = 0.5 contours over (:62-69).

  isotropic_turbulence.npz    test/large_test/isotropic_turbulence/isotropic.basilisk, isotropic.hit3d -- kinetic energy of
                              linearly forced isotropic turbulence (ABC flow + noise, nu = 0.01, forcing 0.1 (v - <v>)) over
                              300 time units from Basilisk (columns t, energy) and from the spectral code hit3d (t, 1.5 x
                              column 2), the two curves isotropic_turbulence/postpro.py:5-17 plots FEN's output.txt over.

Usage (in the build container, where /root/reference is mounted):  python tests/golden/make_reference_data.py
"""
import os

import numpy as np

REF = "/root/reference/test/small_test/multiphase"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    curve = np.genfromtxt(os.path.join(REF, "capillary_wave", "prosperetti.csv"), delimiter=",")
    assert curve.shape == (738, 2)
    np.savez_compressed(os.path.join(HERE, "prosperetti_capillary.npz"), curve=curve)
    com = np.genfromtxt(os.path.join(REF, "rising_bubble", "com_ref.txt"))
    assert com.shape == (2102, 5)
    com2 = np.genfromtxt(os.path.join(REF, "rising_bubble", "com_ref_2.txt"))      # test case 2 (density ratio 1000)
    assert com2.shape == (800, 5)
    shape1 = np.genfromtxt(os.path.join(REF, "rising_bubble", "shape_ref.txt"))     # bubble contour of case 1 at t = 3
    shape1 = shape1[~np.isnan(shape1).any(axis=1)]
    assert shape1.shape == (650, 2)
    np.savez_compressed(os.path.join(HERE, "rising_bubble_com_ref.npz"), t=com[:, 0], yc=com[:, 3], uc=com[:, 4],
                        t2=com2[:, 0], yc2=com2[:, 3], uc2=com2[:, 4], shape1=shape1)
    lid = "/root/reference/test/small_test/navier_stokes/lid_driven"
    uref, vref = np.genfromtxt(os.path.join(lid, "uref")), np.genfromtxt(os.path.join(lid, "vref"))
    assert uref.shape == (17, 2) and vref.shape == (17, 2)
    np.savez_compressed(os.path.join(HERE, "ghia_cavity_re1000.npz"), uref=uref, vref=vref)
    sd = os.path.join(REF, "shear_drop", "reference")
    drops = {"D_Ca%s" % ca: np.genfromtxt(os.path.join(sd, "Re1Ca%sb.csv" % ca), delimiter=",", skip_header=1)
             for ca in ("02", "04", "09")}
    assert drops["D_Ca02"].shape == (11, 2) and drops["D_Ca04"].shape == (21, 2) and drops["D_Ca09"].shape == (31, 2)
    for ca in ("02", "04", "09"):        # Basilisk's drop contours at the end of each case, box-centred coordinates
        drops["shape_Ca%s" % ca] = np.genfromtxt(os.path.join(sd, "shapeRe1Ca%sb.txt" % ca))
    np.savez_compressed(os.path.join(HERE, "shear_drop_basilisk.npz"), **drops)
    iso = "/root/reference/test/large_test/isotropic_turbulence"
    bas, hit = np.genfromtxt(os.path.join(iso, "isotropic.basilisk")), np.genfromtxt(os.path.join(iso, "isotropic.hit3d"))
    assert bas.shape == (12464, 4) and hit.shape == (1438, 4)
    np.savez_compressed(os.path.join(HERE, "isotropic_turbulence.npz"), basilisk=bas[::4, [0, 2]],
                        hit3d=np.stack([hit[:, 0], 1.5 * hit[:, 2]], axis=1))       # postpro.py:15-16
    print("wrote", curve.shape, com.shape, shape1.shape, uref.shape, vref.shape, {k: v.shape for k, v in drops.items()})

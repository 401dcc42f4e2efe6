"""The reference's README / example/example.py quickstart, unchanged, on the B200 implementation.

Run from the repository root (needs a GPU and the built library):
    PYTHONPATH=spart-python_b200 python examples/quickstart.py
"""
import SPART

leafbio = SPART.LeafBiology(Cab=40, Cca=10, Cw=0.02, Cdm=0.01, Cs=0, Cant=10, N=1.5)
soilpar = SPART.SoilParameters(B=0.5, lat=0, lon=100, SMp=20, SMC=25, film=0.015)
canopy = SPART.CanopyStructure(LAI=3, LIDFa=-0.35, LIDFb=-0.15, q=0.05)
angles = SPART.Angles(sol_angle=40, obs_angle=0, rel_angle=0)
atm = SPART.AtmosphericProperties(aot550=0.325, uo3=0.35, uh2o=1.41, Pa=1013.25)

spart = SPART.SPART(soilpar, leafbio, canopy, atm, angles, sensor="Sentinel2A-MSI", DOY=100)
results = spart.run()          # pandas DataFrame: Band, L_TOA, R_TOA, R_TOC indexed by wavelength
print(results)
print("leaf reflectance at 550 / 865 nm:", spart.leafopt.refl[150, 0], spart.leafopt.refl[465, 0])

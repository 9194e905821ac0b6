"""Stage-kernel z-chunk sweep (development tool): python scripts/zchunk_sweep.py [512]; z_chunks = 0 is the library's own choice."""
import sys; sys.path.insert(0,'.')
import numpy as np, breeze_b200 as bz
def run(N, zc):
    grid = bz.RectilinearGrid(bz.B200(use_tma=1, z_chunks=zc), size=(N,N,N), x=(-10e3,10e3), y=(-10e3,10e3), z=(0,10e3))
    m = bz.AtmosphereModel(grid, dynamics=bz.AnelasticDynamics(bz.ReferenceState(grid, potential_temperature=300)))
    m.set(θ=lambda x,y,z: 300+2*np.cos(np.pi/2*np.minimum(1,np.sqrt(x**2+y**2+(z-2000)**2)/2000))**2)
    for _ in range(3): m.time_step(0.5)
    m.context.synchronize(); m.context.profile_enable(True); m.context.profile_read()
    for _ in range(10): m.time_step(0.5)
    ms,n = m.context.profile_read()
    print(f"N={N} z_chunks={zc}: stage {ms[0]/n[0]:.3f} ms/launch, step {ms.sum()/10:.3f} ms", flush=True)
import sys
if len(sys.argv) > 1 and sys.argv[1] == "512":
    for zc in (0, 1, 3, 4, 6): run(512, zc)
else:
    for zc in (0, 2, 4, 8, 11, 16): run(256, zc)
    for zc in (0, 1, 2, 4): run(128, zc)

import sys
sys.path.insert(0, '/root/repo')
import soap_bubble_simulation_b200 as S
for tfe in (False, True):
    for nt, nphi in ((16, 32), (128, 256)):
        a = S.DrainageSimulator(S.SimulationConfig.default(), precision=S.PREC_F32, num_theta=nt, num_phi=nphi)
        a.set_marginal_regeneration(tfe)
        print(nt, nphi, tfe, flush=True)
        a.step(1e-3, 2000); a.sync()

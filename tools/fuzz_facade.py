"""Randomised runs of the drop-in facade on the GPU box: model files -> FileStructure -> BoxData[PM] -> preprocess ->
block_compute[_pm] -> .npz -> load_multiple_results (-> analytic file -> jackknife statistics), 24 random shapes incl.
A_rho != A, one surface, samples that are not a multiple of the block size; the first block of every run against the
oracle on the co-ordinates of the per-block sampler API.

    python tools/fuzz_facade.py
"""
import sys, tempfile, shutil
from os.path import abspath, dirname, join, isfile
ROOT = dirname(dirname(abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pimc_oracle as orc
from pibronic_b200 import file_structure, pimc, synthetic, stats, analytic, constants
from pibronic_b200.model_io import VMK
def rel(a,b): return float(np.max(np.abs(a-b)/np.abs(b)))
rng=np.random.default_rng(5)
bad=0
for k in range(24):
    A=int(rng.integers(1,13)); N=int(rng.choice([1,2,3,6,9,12,24])); P=int(rng.choice([3,8,12,33,64])); pm=bool(rng.integers(0,2))
    T=float(rng.choice([200.0,300.0,500.0])); bs=int(rng.choice([50,100,128])); blocks=int(rng.integers(1,6))
    samples=bs*blocks+int(rng.choice([0,0,7]))
    tmp=tempfile.mkdtemp(prefix="pbx_ff_")
    try:
        FS=file_structure.FileStructure(tmp, k, 0)
        model=synthetic.coupled_model(A,N,(0.05,0.4),(2.0,2.6),seed=100+k,quadratic=float(rng.choice([0.0,0.08])))
        rho=None
        if rng.random()<0.4 and A>1:
            d=synthetic.diagonal_of(model); Ar=int(rng.integers(1,A+3)); reps=-(-Ar//A)
            rho={VMK.N:N,VMK.A:Ar,VMK.w:d[VMK.w],VMK.E:np.tile(d[VMK.E],reps)[:Ar]+0.01*np.arange(Ar),VMK.G1:np.tile(d[VMK.G1],(1,reps))[:,:Ar]*(1+0.05*np.arange(Ar))}
        synthetic.write_data_set(FS, model, rho)
        FS.generate_model_hashes()
        data=(pimc.BoxDataPM if pm else pimc.BoxData).from_FileStructure(FS)
        data.samples,data.beads,data.temperature,data.block_size=samples,P,T,bs
        data.blocks=blocks
        data.hash_vib,data.hash_rho=FS.hash_vib,FS.hash_rho
        data.seed=1000+k
        data.preprocess()
        result=(pimc.BoxResultPM if pm else pimc.BoxResult)(data=data)
        result.path_root,result.id_job=FS.path_rho_results,0
        (pimc.block_compute_pm if pm else pimc.block_compute)(data,result)
        n=bs*blocks
        view=slice(0,bs)
        data.draw_sample(view); data.transform_sampled_coordinates(view)
        R=np.ascontiguousarray(data.qTensor[:,0])
        vib_j=orc.load_vibronic_json(FS.path_vib_model); rho_j=orc.load_sampling_json(FS.path_rho_model)
        tab=orc.precompute(vib_j,rho_j,P,T)
        want=np.stack(orc.estimate_block(tab,R,pm=pm,faithful=False))
        errs=[rel(result.scaled_rho[view],want[0]),rel(result.scaled_g[view],want[1])]
        if pm: errs+= [rel(result.scaled_gofr_plus[view],want[2]),rel(result.scaled_gofr_minus[view],want[3])]
        loaded=type(result)(); loaded.load_multiple_results([result.compute_path_to_file()])
        ok=max(errs)<1e-10 and np.array_equal(loaded.scaled_g[:n],result.scaled_g[:n]) and not np.isnan(result.scaled_rho[:n]).any()
        extra=""
        if pm and n==samples:
            analytic.analytic_of_sampling_model(FS, constants.beta(T))
            try:
                stats.jackknife_analysis_of_pimc(FS, method="basic"); extra="stats ok"
            except Exception as e:
                extra="stats: %r"%(e,); ok=False
        print(f"[{k:2d}] A={A} N={N} Ar={data.rho.states if hasattr(data.rho,'states') else '?'} P={P} pm={int(pm)} T={T} n={n}/{samples}: max err {max(errs):.1e} {extra} {'ok' if ok else 'FAIL'}",flush=True)
        bad+=not ok
        data.release()
    finally:
        shutil.rmtree(tmp,ignore_errors=True)
print("failures",bad)

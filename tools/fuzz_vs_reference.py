"""Differential fuzz of the oracle against the REFERENCE'S OWN FILES (build container only: needs /root/reference).

    python tools/fuzz_vs_reference.py            # ~1 min; prints every mismatch, exits 1 if there is one

Part 1: 24 random cases over 6 grid shapes (odd sizes included), shifted domain origins, three boundary types with random
wall values, field scales 0.01 / 1 / 30, a zero component, dt from 1e-4 to 2e-2: the three advection schemes, laplacian
and divergence must be BIT-IDENTICAL, the three pressure solves within 2e-5.  Part 2: 10 random immersed-body cases (1-3
bodies anywhere, including across the domain edge; anisotropic cells; random harmonic laws and time stamps): the oracle's
dense and windowed forces against the reference's dense sums, bar 1e-5.  Part 3: 12 random penalty masks (ellipse tables,
centres, smoothing constants) and 12 x 300 tracer sample points (inside, on the edges, outside; shifted origins): bit-identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
sys.path.insert(0, ROOT)
import numpy as np
import make_reference_golden as m
R=m.load_reference()
import jax.numpy as jnp
from oracle import field, ib, kinematics as okin, poisson, stencils
F=np.float32
rng=np.random.default_rng(7)
bad=0
for trial in range(24):
    nx,ny=[(33,20),(40,72),(16,16),(21,35),(64,30),(12,50)][trial%6]
    dom=((float(rng.uniform(-1,1)),0),(float(rng.uniform(-1,1)),0)); dom=((dom[0][0],dom[0][0]+rng.uniform(0.5,3)),(dom[1][0],dom[1][0]+rng.uniform(0.5,3)))
    grid=R.grids.Grid((nx,ny),domain=dom); og=field.Grid((nx,ny),dom)
    scale=rng.choice([0.01,1.0,30.0])
    u,v,p=((scale*rng.standard_normal((nx,ny))).astype(F) for _ in range(3))
    if trial%5==0: u[:]=0
    dt=float(rng.choice([1e-4,1e-3,2e-2]))
    w=rng.uniform(-1,1,(2,2,2))
    kind=["periodic","channel","far_field"][trial%3]
    if kind=="periodic":
        rb=(R.boundaries.periodic_boundary_conditions(2),)*2; ob=(field.periodic_bc(),)*2
    elif kind=="channel":
        vals=[((0.0,0.0),(float(w[k,1,0]),float(w[k,1,1]))) for k in range(2)]
        rb=tuple(R.boundaries.Moving_wall_boundary_conditions(2,vals[k],0.0,None) for k in range(2)); ob=tuple(field.channel_bc(vals[k]) for k in range(2))
    else:
        vals=[tuple(tuple(float(x) for x in r) for r in w[k]) for k in range(2)]
        rb=tuple(R.boundaries.Far_field_boundary_conditions(2,vals[k],0.0,None) for k in range(2)); ob=tuple(field.far_field_bc(vals[k]) for k in range(2))
    U=R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u),(1.0,.5),grid),rb[0]); V=R.grids.GridVariable(R.grids.GridArray(jnp.asarray(v),(.5,1.0),grid),rb[1])
    oU=field.Var(u,(1.0,.5),og,ob[0]); oV=field.Var(v,(.5,1.0),og,ob[1])
    with np.errstate(all="ignore"):
        for name,rf,of in (("upwind",R.advection.advect_upwind,stencils.advect_upwind),("vl",R.advection.advect_van_leer,stencils.advect_van_leer),("tvd",R.advection.advect_van_leer_using_limiters,stencils.advect_van_leer_using_limiters)):
            for k,(c,oc) in enumerate(((U,oU),(V,oV))):
                a=np.asarray(rf(c,(U,V),dt).data); b=of(oc,(oU,oV),dt)
                if not np.array_equal(a,b,equal_nan=True):
                    bad+=1; print("MISMATCH",trial,kind,(nx,ny),name,k,np.nanmax(np.abs(a-b)))
        a=np.asarray(R.finite_differences.laplacian(U).data); b=stencils.laplacian(oU)
        if not np.array_equal(a,b): bad+=1; print("MISMATCH lap",trial)
        a=np.asarray(R.finite_differences.divergence((U,V)).data); b=stencils.divergence((oU,oV))
        if not np.array_equal(a,b): bad+=1; print("MISMATCH div",trial)
        solver={"periodic":R.pressure.solve_fast_diag,"channel":R.pressure.solve_fast_diag_moving_wall,"far_field":R.pressure.solve_fast_diag_Far_Field}[kind]
        q=np.asarray(solver((U,V)).data); oq=poisson.SOLVERS[kind]((oU,oV))
        e=np.linalg.norm(q.astype(float)-oq)/max(np.linalg.norm(oq),1e-30)
        if not e<2e-5: bad+=1; print("MISMATCH solve",trial,kind,(nx,ny),e)
print("stencils and solves: mismatches:", bad)

# ---------------------------------------------------------------- part 2: immersed bodies
delta=R.convolution_functions.delta_approx_logistjax
surf=lambda f,xp,yp: R.convolution_functions.new_surf_fn(f,xp,yp,delta)
worst=0
rng=np.random.default_rng(11)
for trial in range(10):
    nx,ny=[(40,36),(33,48),(50,50)][trial%3]
    x0d,y0d=float(rng.uniform(-1,1)),float(rng.uniform(-1,1))
    h=float(rng.uniform(0.03,0.06)); hy=h*float(rng.choice([1.0,0.8,1.25]))
    dom=((x0d,x0d+nx*h),(y0d,y0d+ny*hy))
    grid=R.grids.Grid((nx,ny),domain=dom); og=field.Grid((nx,ny),dom)
    u,v=((rng.standard_normal((nx,ny))).astype(F) for _ in range(2))
    t0=F(rng.uniform(0,3)); dt=1e-3
    per=R.boundaries.periodic_boundary_conditions(2)
    bc=R.boundaries.ConstantBoundaryConditions(values=((0.,0.),(0.,0.)),time_stamp=t0,types=per.types,boundary_fn=None)
    U=R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u),(1.0,.5),grid),bc); V=R.grids.GridVariable(R.grids.GridArray(jnp.asarray(v),(.5,1.0),grid),bc)
    P=R.grids.GridVariable(R.grids.GridArray(jnp.zeros((nx,ny)),(.5,.5),grid),per)
    nb=int(rng.integers(1,4))
    # centres anywhere, including next to / across the domain edge (no periodic image in the reference)
    centers=np.stack([rng.uniform(dom[0][0]-0.1,dom[0][1]+0.1,nb),rng.uniform(dom[1][0]-0.1,dom[1][1]+0.1,nb)],1).astype(F)
    a=float(rng.uniform(0.1,0.4)); n=int(rng.integers(12,40))
    xm,ym=m.ellipse_markers(a,0.3*a,n)
    dpar=[[float(rng.uniform(0,1)),float(rng.uniform(0.1,1))] for _ in range(nb)]
    rpar=[[float(rng.uniform(0,3)),float(rng.uniform(0,1)),float(rng.uniform(0.1,1)),float(rng.uniform(0,3))] for _ in range(nb)]
    shape_fn=lambda g,gp:(jnp.asarray(xm),jnp.asarray(ym))
    g1=R.particle_class.Grid1d(len(xm),domain=(0,2*np.pi))
    parts=R.particle_class.particle(jnp.asarray(centers),[[a,0.3*a]]*nb,dpar,rpar,g1,shape_fn,R.kinematics.displacement,R.kinematics.rotation)
    av=R.particle_class.All_Variables(parts,(U,V),P,[],0,None)
    fx,fy=R.IBM_Force.calc_IBM_force_NEW_MULTIPLE(av,delta,surf,dt)
    ob=ib.Bodies(centers,[[a,0.3*a]]*nb,dpar,rpar,lambda g:(xm,ym),okin.DISPLACEMENT["harmonic"],okin.ROTATION["harmonic"])
    oU=field.Var(u,(1.0,.5),og,field.periodic_bc(t0)); oV=field.Var(v,(.5,1.0),og,field.periodic_bc(t0))
    ofx,ofy=ib.ibm_force((oU,oV),ob,dt,R=None)
    rel=lambda a_,b_: np.linalg.norm(a_.astype(float)-b_)/max(np.linalg.norm(b_),1e-30)
    e=max(rel(ofx,np.asarray(fx.data)),rel(ofy,np.asarray(fy.data)))
    # windowed oracle (what the kernels are compared with) against the reference's dense sums
    R_eff=int(np.ceil(6*max(1.0,h/hy)-1e-9))
    wfx,wfy=ib.ibm_force((oU,oV),ob,dt,R=R_eff)
    ew=max(rel(wfx,np.asarray(fx.data)),rel(wfy,np.asarray(fy.data)))
    worst=max(worst,e,ew)
    print(trial,(nx,ny),nb,"dense",e,"windowed R=%d"%R_eff,ew)
print("immersed bodies: worst rel-L2", worst)

# ---------------------------------------------------------------- part 3: penalty mask, tracer sampling (bit-identical)
from oracle import penalty, tracers
rng=np.random.default_rng(5)
bad3=0
for trial in range(12):
    nx,ny=[(40,36),(33,48),(64,64)][trial%3]
    dom=((float(rng.uniform(-1,1)),),(float(rng.uniform(-1,1)),)); dom=((dom[0][0],dom[0][0]+rng.uniform(1,3)),(dom[1][0],dom[1][0]+rng.uniform(1,3)))
    grid=R.grids.Grid((nx,ny),domain=dom); og=field.Grid((nx,ny),dom)
    n=int(rng.integers(20,90)); theta=np.linspace(0,2*np.pi,n).astype(F)
    a,b=float(rng.uniform(0.2,0.6)),float(rng.uniform(0.1,0.3))
    Rth=np.asarray(R.parameteric_fns.param_ellipse([a,b],jnp.asarray(theta)))
    c=(float(rng.uniform(*dom[0])),float(rng.uniform(*dom[1])))
    know=(float(rng.uniform(100,5000)),float(rng.uniform(0.005,0.05)))
    smooth=lambda G_,K: K[0]/2.0*(1.0+jnp.tanh(G_/K[1]))
    with np.errstate(all="ignore"):
        ref=np.asarray(R.util_funs.calc_perm(grid,c,jnp.asarray(Rth),smooth,know))
        got=penalty.calc_perm(og,c,Rth,penalty.tanh_smoothing,know)
    if not np.array_equal(ref,got,equal_nan=True):
        bad3+=1; print("perm mismatch",trial,np.nanmax(np.abs(ref-got)))
    u=rng.standard_normal((nx,ny)).astype(F)
    U=R.grids.GridVariable(R.grids.GridArray(jnp.asarray(u),(1.0,.5),grid),R.boundaries.periodic_boundary_conditions(2))
    pts=np.stack([rng.uniform(dom[0][0]-0.2,dom[0][1]+0.2,300),rng.uniform(dom[1][0]-0.2,dom[1][1]+0.2,300)],1).astype(F)
    ref=np.asarray(R.CFD_MD_coupling.interpolate_pbc(U,jnp.asarray(pts)))
    got=tracers.interpolate_field(field.Var(u,(1.0,.5),og,field.periodic_bc()),pts)
    if not np.array_equal(ref,got):
        bad3+=1; print("sample mismatch",trial,np.abs(ref-got).max())
print("penalty mask and tracer sampling: mismatches:", bad3)
sys.exit(1 if (bad or bad3 or not worst < 1e-5) else 0)

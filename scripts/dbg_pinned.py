import numpy as np, sys
sys.path.insert(0,'.')
from bench import hot_params
from tensormol_b200.SystemBuilders import water_box, wrap_into_cell
from tensormol_b200.engine import Engine, random_weights
Z,X,lat=water_box(6,spacing=3.1072,seed=3); X=wrap_into_cell(X,lat)
eng=Engine([1,8],[64,64,64],hot_params()); eng.set_weights(random_weights([1,8],eng.D,[64,64,64],0))
ref=eng.evaluate_lattice(X,Z,lat,1)
Xp,Zp=eng.pinned(X.shape),eng.pinned(Z.shape,np.int32); Xp[:]=X; Zp[:]=Z
into={"gradient":eng.pinned((1,len(Z),3)),"charge":eng.pinned((1,len(Z)))}
for i in range(6):
    r=eng.evaluate_lattice(Xp,Zp,lat,1,into=into)
    assert r["gradient"] is into["gradient"]
    print(i, abs(r["gradient"]-ref["gradient"]).max(), abs(r["charge"]-ref["charge"]).max(), abs(r["Etotal"]-ref["Etotal"]).max(), abs(r["Ebp_atom"]-ref["Ebp_atom"]).max())
    into["gradient"][:]=0; into["charge"][:]=7
# moved coordinates in the same pinned array must be picked up by the replay
Xp[:]=wrap_into_cell(X+0.05*np.random.default_rng(0).normal(size=X.shape),lat)
r2=eng.evaluate_lattice(Xp,Zp,lat,1,into=into); ref2=eng.evaluate_lattice(np.array(Xp),np.array(Zp),lat,1)
print("moved", abs(r2["gradient"]-ref2["gradient"]).max(), abs(r2["Etotal"]-ref2["Etotal"]).max(), abs(ref2["Etotal"]-ref["Etotal"]).max())

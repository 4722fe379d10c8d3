import cProfile, pstats, sys, time
import numpy as np
sys.path.insert(0, ".")
import varnet_b200
from varnet_b200 import ManageTrainData
pi = np.pi
def diffFun(x, t=0, D=0.1 / pi):
    return D * np.ones([np.shape(x)[0], 1])
mor = varnet_b200.MOR(diffFun, ['D'], [[0.003, 0.033]])
pde = varnet_b200.ADPDE(varnet_b200.Domain1D(), diff=diffFun, vel=1.0, timeDependent=True, tInterval=[0, 2.0], IC=lambda x: -np.sin(pi * x), MORvar=mor)
disc_f = lambda n=6: np.array([0.003 * (11 ** (k / (n - 1))) for k in range(n)])[np.newaxis].T
vn = varnet_b200.VarNet(pde, layerWidth=[10, 20, 30], discNum=150, bDiscNum=75, tDiscNum=800, MORdiscScheme=disc_f, processors='GPU:0', seed=0)
fd = vn.fixData; fd.setFEdata()
Input, _, biInput, _ = vn.trainingPoints()
disc = vn.PDE.MORvar.discretizeArg(vn.MORdiscScheme)
tData = ManageTrainData(Input, biInput, 20, None, True, fd.MORbatchNum)
tData = vn.trainData(0, disc, tData)
tData.updateDictFields('trainW', np.array([10., 10., 1.]))
tf = vn.tfData
def epoch():
    global tData
    total = 0
    for b in range(fd.MORbatchNum):
        tData = vn.trainData(b, disc, tData)
        total += tData.optimIter(tf)
    return float(total)
epoch(); epoch()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(20):
        loss = epoch()
    dt = (time.perf_counter() - t0) / 20
    print("epoch %.3f ms  %.0f steps/s  loss %.6g  uploads %d  defer %s" % (dt * 1e3, fd.MORbatchNum * tData.batchNum / dt, loss, tf.uploads, tf.defer_losses))
cProfile.run("epoch()", "/tmp/prof.out")
pstats.Stats("/tmp/prof.out").sort_stats("cumulative").print_stats(18)

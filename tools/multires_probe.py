"""Forward / gradient accuracy of the low-resolution observation at cfg4 size, aligned and rotated (GPU, float32 plan vs oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scarlet_b200 import synthetic
from oracle import scenes


def rel_peak(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


for name, cfg in (("aligned", dict(synthetic.CFG4)), ("rotated", dict(synthetic.CFG4, lr_angle=25.0, config_id=41))):
    scene = synthetic.make_multires_scene(0, cfg)
    blend = synthetic.make_multires_blend(scene, precision=32)
    o = scenes.build_multires_oracle(scene, scenes.multires_setup(blend))
    plan = blend._get_plan()
    plan.upload_parameters(state=False)
    model = o.get_model()
    ev0 = plan.evaluate(obs=0, want=("model", "rendered", "loss", "grads"))
    loss, grads = o.loss_and_grads()
    gscale = max(np.abs(gr).max() for gr in grads[1::3])
    print(name, "model", rel_peak(ev0["model"][0], model), "rendered", rel_peak(ev0["rendered"][0], o.observations[0].render(model)),
          "loss", abs(ev0["loss"][0] - loss) / abs(loss),
          "g_sed", max(rel_peak(ev0["g_sed"][k], grads[3 * k]) for k in range(len(o.sources))),
          "g_morph", max(np.abs(ev0["g_morph"][k] - grads[3 * k + 1]).max() for k in range(len(o.sources))) / gscale)

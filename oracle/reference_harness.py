"""Oracle tooling (build container only): run the UNMODIFIED reference from /root/reference on CPU and record what the
golden fixtures need.  Nothing here travels to the GPU box; only its outputs (tests/golden/*.pt) do.

What is patched, and why:
  * the path hard-codes ``.cuda()`` and ``validate()`` returns None without CUDA  -> 3-line shim;
  * ``iter_num = 8`` is a literal (language_eval.py:136)                          -> replaced, in an in-memory copy of
    the module source, by ``iter_num = opt.neval_episodes`` so short runs are possible;
  * recording hooks wrap ``criterion``, ``net.regloss``, ``net.reglossnovel``, ``LangPuller.loss1`` and ``validate``;
    they only observe values (the loss total is re-assembled with the same fp32 additions, in the same order).
"""
import contextlib
import importlib
import sys
import types

import numpy as np
import torch

REF = '/root/reference'


@contextlib.contextmanager
def reference_modules():
    """Import the reference's packages (models, eval, dataset) with the CPU shim; restore sys.modules afterwards."""
    saved_mods = {k: v for k, v in sys.modules.items() if k.split('.')[0] in ('models', 'eval', 'dataset', 'util', 'configs')}
    for k in list(saved_mods):
        del sys.modules[k]
    saved = (torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.is_available)
    saved_capturing = torch.cuda.is_current_stream_capturing
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    # torch.optim.Adam.step asks whether a CUDA graph is being captured whenever "CUDA is available" (the shim above)
    torch.cuda.is_current_stream_capturing = lambda: False
    sys.path.insert(0, REF)
    try:
        models_util = importlib.import_module('models.util')
        resnet_language = importlib.import_module('models.resnet_language')
        src = open(REF + '/eval/language_eval.py').read()
        assert src.count("iter_num = 8") == 1
        src = src.replace("iter_num = 8", "iter_num = opt.neval_episodes")
        importlib.import_module('eval')                       # namespace package of the reference
        le = types.ModuleType('eval.language_eval')
        le.__package__ = 'eval'
        le.__file__ = REF + '/eval/language_eval.py'
        sys.modules['eval.language_eval'] = le
        exec(compile(src, le.__file__, 'exec'), le.__dict__)
        yield types.SimpleNamespace(models_util=models_util, resnet_language=resnet_language, language_eval=le)
    finally:
        sys.path.remove(REF)
        torch.Tensor.cuda, torch.nn.Module.cuda, torch.cuda.is_available = saved
        torch.cuda.is_current_stream_capturing = saved_capturing
        for k in [k for k in sys.modules if k.split('.')[0] in ('models', 'eval', 'dataset', 'util', 'configs')]:
            del sys.modules[k]
        sys.modules.update(saved_mods)


def run_reference(world, n_sessions, seed, probe_rows=0):
    """-> (record, initial_state_dict).  Record layout matches oracle.session.run_sessions."""
    from srb200 import synthetic
    opt = world.opt
    opt.neval_episodes = n_sessions
    with reference_modules() as ref:
        net = synthetic.init_model(ref.models_util.create_model, opt, seed)
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        ckpt = synthetic.make_ckpt(net, world)
        rec = dict(sessions=[], timers=None)
        cur = dict(terms=[], row=None, W=[], preds=None)

        ce = torch.nn.CrossEntropyLoss()
        calls = dict(ce=[])

        def criterion(out, y):
            v = ce(out, y)
            if torch.is_grad_enabled():
                calls['ce'].append(v.item())
            return v
        orig_regloss, orig_reglossnovel = net.regloss, net.reglossnovel
        vals = dict(b=None, n=None, p=None)

        def regloss(*a, **k):
            r = orig_regloss(*a, **k)
            vals['b'] = r.item()
            return r

        def reglossnovel(*a, **k):
            r = orig_reglossnovel(*a, **k)
            vals['n'] = r.item()
            return r
        net.regloss, net.reglossnovel = regloss, reglossnovel
        LP = ref.resnet_language.LangPuller
        orig_loss1 = LP.loss1

        def loss1(self, *a, **k):
            r = orig_loss1(self, *a, **k)
            vals['p'] = r.item()
            return r
        LP.loss1 = loss1
        le = ref.language_eval
        orig_validate = le.validate

        def validate(query_xs, query_ys_id, net_, criterion_, opt_, epoch):
            r = orig_validate(query_xs, query_ys_id, net_, criterion_, opt_, epoch)
            if isinstance(query_xs, list):      # top-level call: one epoch finished
                f32 = np.float32
                ce_s = calls['ce'][0]
                ce_m = calls['ce'][1] if len(calls['ce']) > 1 else None
                tot = f32(ce_s)
                for t in (ce_m, vals['b'], vals['n'], vals['p']):
                    if t is not None:
                        tot = f32(tot + f32(t))
                cur['terms'].append([float(tot), ce_s, ce_m or 0.0, vals['b'] or 0.0, vals['n'] or 0.0, vals['p'] or 0.0])
                cur['W'].append(net_.classifier.weight.detach().clone())
                cur['last'] = r
                calls['ce'] = []
                vals.update(b=None, n=None, p=None)
            return r
        le.validate = validate
        orig_log = le.log_episode

        def log_episode(novel_labels, vocab_novel, epoch, novel_acc, base_acc, running_base, running_novel):
            last = cur['last']
            srec = dict(epochs=epoch - 1, terms=np.asarray(cur['terms'], dtype=np.float64), W=cur['W'][-1].clone(),
                        W_traj=torch.stack(cur['W']), novel_session_acc=[round(i.item(), 2) for i in last[0]],
                        query_pred=[torch.from_numpy(np.asarray(p)).long() for p in last[3]], acc_base=base_acc,
                        vocab_novel=list(vocab_novel))
            srec['bn'] = {k: v.clone() for k, v in net.state_dict().items() if 'running_' in k or 'num_batches_tracked' in k}
            if probe_rows:
                net.eval()
                with torch.no_grad():
                    saved = {n: m.num_batches_tracked for n, m in net.named_modules() if isinstance(m.__dict__.get('num_batches_tracked', None), int)}
                    _, x = None, world.meta_valloader.batches[len(rec['sessions'])][0]
                    x = x.view(-1, *x.shape[2:])[:probe_rows]
                    feats, _ = net(x, is_feat=True)
                    srec['probe_feat'] = feats[-1].clone()
                    for n, m in net.named_modules():
                        if n in saved:
                            m.num_batches_tracked = saved[n]
            rec['sessions'].append(srec)
            cur.update(terms=[], W=[])
            return orig_log(novel_labels, vocab_novel, epoch, novel_acc, base_acc, running_base, running_novel)
        le.log_episode = log_episode

        import io
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            novel_avg, base_avg = le.few_shot_finetune_incremental_test(
                net, ckpt, criterion, world.meta_valloader, world.base_val_loader, opt,
                base_support_loader=world.base_support_loader)
        out = buf.getvalue()
        rec['acc_novel_avg'], rec['acc_base_avg'] = float(novel_avg), float(base_avg)
        for line in out.splitlines():
            for key, name in (("Overall continual accuracies:", 'weighted'), ("Novel only incremental:", 'novel'),
                              ("Base only incremental:", 'base')):
                if line.startswith(key):
                    rec[name] = [float(x) for x in eval(line[len(key):].replace('np.float64', ''))]
        rec['stdout_tail'] = out[-2000:]
        rec['counters'] = {n: m.num_batches_tracked for n, m in net.named_modules()
                           if isinstance(m.__dict__.get('num_batches_tracked', None), int)}
        rec['final_state'] = {k: v.clone() for k, v in net.state_dict().items()}
        LP.loss1 = orig_loss1
    return rec, sd0

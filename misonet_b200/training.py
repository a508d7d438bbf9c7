"""MISO_1 training step (mirror of ``Trainer_Separate._run_one_epoch``'s batch body, trainer.py:146-212):
roll the reference microphone to the front, ``estimate = model(mix)``, ``loss_uPIT`` against the clean images at the
reference microphone, ``loss.backward()``, gradient clipping, optimizer step.  The arithmetic (forward, loss, backward)
runs in the CUDA library; the optimizer is the caller's ``torch.optim`` object over the module's parameters, exactly as
in run.py, and multi-GPU training adds ONE collective: the gradient all-reduce (distributed.allreduce_gradients)."""
import torch

from . import criterion, distributed


def _backward(model, loss, n_local):
    """loss.backward() plus the data-parallel gradient reduction.  The library's modules reduce their flat gradient
    buffer in place inside backward (``data_parallel``, model._NetFunction); any other module gets the generic bucketed
    all-reduce."""
    if hasattr(model, "data_parallel"):
        model.data_parallel = True
        loss.backward()
    else:
        loss.backward()
        distributed.allreduce_gradients(model.parameters(), n_local=n_local)


def train_step(model, optimizer, mix_stft, ref_stft, ref_ch=0, max_norm=None, training=True):
    """mix_stft: complex [B, Mic, T, F] (CUDA); ref_stft: list[num_spks] of complex [B, Mic, T, F] or [B, T, F]
    (the clean source images; trainer.py:160-166 takes microphone ``ref_ch``).  Returns the loss (float32 CUDA scalar,
    detached).  With ``training=False`` this is the validation pass of trainer.py:138-144 (no_grad, no update)."""
    num_spks = model.num_spks
    if hasattr(model, "forward_shifts"):
        # trainer.py:154 rolls the reference microphone to the front; the library's pack kernel applies the circular
        # shift while it writes the input planes, so no rolled copy of the mixture is made
        mix, net = mix_stft, (lambda x: model.forward_shifts(x, (ref_ch,)))
    else:
        mix, net = torch.roll(mix_stft, -ref_ch, dims=1), model
    refs = [r[:, ref_ch] if r.dim() == 4 else r for r in ref_stft]               # trainer.py:162-166
    if not training:
        with torch.no_grad():
            return criterion.loss_uPIT(num_spks, net(mix), refs)
    optimizer.zero_grad(set_to_none=True)
    estimate = net(mix)                                                          # trainer.py:158
    if estimate.shape[1] != num_spks:
        raise ValueError("[ERROR] please check the number of speakers")          # trainer.py:169
    loss = criterion.loss_uPIT(num_spks, estimate, refs)                         # trainer.py:172
    _backward(model, loss, mix.shape[0])                                         # trainer.py:207
    if max_norm:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)             # trainer.py:209-211
    optimizer.step()                                                             # trainer.py:212
    return loss.detach()


def train_step_enhance(model, optimizer, mix_stft, beamform_stft, miso1_stft, ref_stft, max_norm=None, training=True):
    """MISO_3 enhancement step (mirror of ``Trainer_Enhance``'s batch body, trainer.py:388-416): one optimizer update per
    speaker, ``model(mix, s_bf, MISO1_spk)`` against ``loss_Enhance``.  ``beamform_stft`` / ``miso1_stft`` / ``ref_stft``:
    list[num_spks] of complex [B, 1, T, F] (beamformed, MISO1 at the reference microphone, clean reference; they are
    data -- trainer.py reads them from the pickles of data.py:133-207 -- so only the parameters receive gradients).
    Returns the mean of the per-speaker losses (trainer.py:410,423 accumulate ``loss.item() / 2``).

    trainer.py:416 feeds speaker 1's beamformed signal to speaker 2's training forward (SURVEY.md appendix B lists it as
    a defect; the inference path tester.py:936-939 and the validation branch trainer.py:413 use speaker 2's); this
    mirror pairs each speaker with its own beamformed signal."""
    total = 0.0
    n = len(ref_stft)
    for s in range(n):
        if not training:
            with torch.no_grad():
                total = total + criterion.loss_Enhance(model(mix_stft, beamform_stft[s], miso1_stft[s]), ref_stft[s]) / n
            continue
        estimate = model(mix_stft, beamform_stft[s], miso1_stft[s])              # trainer.py:400,416
        loss = criterion.loss_Enhance(estimate, ref_stft[s])                     # trainer.py:402,418
        optimizer.zero_grad(set_to_none=True)
        _backward(model, loss, mix_stft.shape[0])
        if max_norm:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()
        total = total + loss.detach() / n
    return total

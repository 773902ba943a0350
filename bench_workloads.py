"""Secondary workloads of BASELINE.json (configs[2..4]) measured with the same rules as bench.py:
CUDA-event timing after warm-up, synthetic ids / images, random-init weights.

  transformer : tests/transformer.ini at the perf shape - 6 layers, d=512, 8 heads, F=2048,
                V=32000, 4096 target tokens per step (64 x 64), Adam; train tokens/s
  beam        : tests/beamsearch.ini at the perf shape - beam 8 over the 6x512 Transformer,
                forced full-length hypotheses (EOS suppressed); emitted tokens/s and batch-1 latency
  captioning  : tests/captioning.ini - frozen VGG-16 conv stack on 224x224 images + Bahdanau
                decoder; images/s of a training step

Usage: python bench_workloads.py [transformer|beam|captioning] ...   (one JSON line each)
"""
import json
import sys
import time

import torch

SEED = 2574600


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time(fn, steps, warmup):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def build_transformer(vocab=32000, dim=512, ff=2048, depth=6, heads=8, max_len=64, lr=1e-4, tie=True):
    from neuralmonkey_b200 import runtime, tf
    from neuralmonkey_b200.decoders import TransformerDecoder
    from neuralmonkey_b200.encoders import TransformerEncoder
    from neuralmonkey_b200.model.sequence import EmbeddedSequence
    from neuralmonkey_b200.trainers import CrossEntropyTrainer
    from neuralmonkey_b200.vocabulary import Vocabulary

    runtime.reset()
    words = ["w{}".format(i) for i in range(vocab - 4)]
    src_vocab, tgt_vocab = Vocabulary(words), Vocabulary(words)
    seq = EmbeddedSequence(name="input_sequence", vocabulary=src_vocab, data_id="source",
                           embedding_size=dim, max_length=max_len, scale_embeddings_by_depth=True)
    enc = TransformerEncoder(name="encoder", input_sequence=seq, ff_hidden_size=ff, depth=depth,
                             n_heads=heads)
    dec = TransformerDecoder(name="decoder", encoders=[enc], vocabulary=tgt_vocab, data_id="target",
                             ff_hidden_size=ff, n_heads_self=heads, n_heads_enc=heads, depth=depth,
                             max_output_len=max_len, embedding_size=dim, tie_embeddings=tie)
    trainer = CrossEntropyTrainer(decoders=[dec], optimizer=tf.AdamOptimizer(learning_rate=lr), use_cuda_graph=True)
    for part in trainer.parameterizeds:
        part.ensure_declared()
    runtime.arena().finalize(runtime.device())
    return {"seq": seq, "enc": enc, "dec": dec, "trainer": trainer, "arena": runtime.arena()}


def feed_transformer(model, src, tgt, train):
    bsz = src.shape[0]
    model["seq"].feed_ids([src], train=train)
    enc = model["enc"]
    enc.reset_batch()
    enc.train_mode = train
    enc.batch_size = bsz
    model["dec"].feed_ids(tgt, bsz, train=train)


def synthetic_ids(bsz, length, vocab, seed, eos=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(4, vocab, (bsz, length), generator=g)
    if eos:
        ids[:, -1] = 2
    return ids


def run_transformer(steps=5, warmup=3, bsz=64, length=64, vocab=32000, breakdown=False):
    model = build_transformer(vocab=vocab, max_len=length)
    batches = [(synthetic_ids(bsz, length, vocab, SEED + i, eos=False).pin_memory(),
                synthetic_ids(bsz, length, vocab, SEED + 100 + i).pin_memory()) for i in range(4)]

    def step(i):
        src, tgt = batches[i % len(batches)]
        feed_transformer(model, src, tgt, train=True)
        return model["trainer"].train_step()

    ms = _time(step, steps, warmup)
    loss = float(step(0)["losses"][0])
    if breakdown:
        from neuralmonkey_b200 import lib
        lib.profile_start()
        step(1)
        prof = lib.profile_stop()
        table = sorted(((n, d["ms"], d["calls"]) for n, d in prof.items()), key=lambda x: -x[1])
        for n, t, c in table[:30]:
            print("# {:34s} {:8.3f} ms  {:4d} calls".format(n, t, c), file=sys.stderr)
        print("# sum {:.3f} ms vs {:.3f} ms per step".format(sum(t for _, t, _ in table), ms), file=sys.stderr)
    return {"workload": "tests/transformer.ini perf shape: L=6 d=512 h=8 F=2048 V={} batch {}x{}".format(
                vocab, bsz, length),
            "metric": "train_target_tokens_per_sec", "value": bsz * length / (ms * 1e-3),
            "unit": "tokens/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "last_loss": loss,
            "params": int(model["arena"].trainable_size), "inputs": "pinned host ids, H2D inside the timed region"}


def run_beam(bsz=64, beam=8, steps=128, vocab=32000, src_len=32, reps=1):
    from neuralmonkey_b200.decoders import BeamSearchDecoder
    model = build_transformer(vocab=vocab, max_len=max(steps, src_len), tie=False)
    dec = model["dec"]
    # EOS never wins: every hypothesis runs the full length (SURVEY 8(d) "forced steps" variant)
    with torch.no_grad():
        dec.var("state_to_word_b")[2] = -1.0e4
    bs = BeamSearchDecoder(name="bs", parent_decoder=dec, beam_size=beam, max_steps=steps,
                           length_normalization=0.6)
    out = {}
    for label, b in (("batch", bsz), ("latency", 1)):
        src = synthetic_ids(b, src_len, vocab, SEED, eos=False)

        def run(_i, b=b, src=src):
            feed_transformer(model, src, None, train=False)
            bs.reset_batch()
            bs.batch_size = b
            return bs.outputs

        for _ in range(3):       # loop run, capture run, first replay
            run(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = _events()
        e0.record()
        for i in range(reps):
            res = run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        emitted = int(res.last_search_step_output.token_ids.shape[0] - 1)
        out[label] = {"batch": b, "ms_per_batch": ms, "steps_run": emitted,
                      "tokens_per_s": b * emitted / (ms * 1e-3),
                      "wall_ms": (time.perf_counter() - t0) * 1e3 / reps}
    return {"workload": "tests/beamsearch.ini perf shape: beam {} x {} steps over the 6x512 Transformer, "
                        "V={}, EOS suppressed".format(beam, steps, vocab),
            "metric": "beam_decode_emitted_tokens_per_sec", "value": out["batch"]["tokens_per_s"],
            "unit": "tokens/s", "batch_throughput": out["batch"], "batch1_latency": out["latency"],
            "note": "self-attention keys/values of the prefix are cached per hypothesis and re-ordered "
                    "with the beam (the reference re-runs the whole prefix every step: "
                    "decoders/transformer.py:485-518)"}


def run_captioning(steps=3, warmup=2, bsz=32, vt=10000, ty=16):
    from neuralmonkey_b200 import runtime, tf
    from neuralmonkey_b200.attention import Attention
    from neuralmonkey_b200.decoders import Decoder
    from neuralmonkey_b200.encoders import ImageNet
    from neuralmonkey_b200.trainers import CrossEntropyTrainer
    from neuralmonkey_b200.vocabulary import Vocabulary

    runtime.reset()
    vocab = Vocabulary(["t{}".format(i) for i in range(vt - 4)])
    enc = ImageNet(name="imagenet_vgg", data_id="images", network_type="vgg_16",
                   spatial_layer="vgg_16/conv5/conv5_3")
    att = Attention(name="attention", encoder=enc, state_size=512)
    dec = Decoder(encoders=[enc], vocabulary=vocab, data_id="target", name="decoder", max_output_len=ty,
                  rnn_size=512, embedding_size=512, attentions=[att])
    trainer = CrossEntropyTrainer(decoders=[dec], optimizer=tf.AdamOptimizer(learning_rate=1e-4), use_cuda_graph=True)
    for part in trainer.parameterizeds:
        part.ensure_declared()
    runtime.arena().finalize(runtime.device())
    g = torch.Generator().manual_seed(SEED)
    images = (torch.randn(bsz, 224, 224, 3, generator=g) * 60.0).pin_memory()
    tgt = synthetic_ids(bsz, ty, vt, SEED + 1)

    def step(_i):
        enc.feed_images(images, train=True)
        att.reset_batch()
        att.train_mode, att.batch_size = True, bsz
        dec.feed_ids(tgt, bsz, train=True)
        return trainer.train_step()

    ms = _time(step, steps, warmup)
    return {"workload": "tests/captioning.ini scaled: VGG-16 conv5_3 on {}x224x224x3 + GRU-512 Bahdanau "
                        "decoder, V={}".format(bsz, vt),
            "metric": "train_images_per_sec", "value": bsz / (ms * 1e-3), "unit": "images/s",
            "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "conv_gflop_per_image": 30.7,
            "conv_tflops": bsz * 30.7e9 / (ms * 1e-3) / 1e12,
            "note": "convolutions = im2col + tcgen05 GEMM (TF32) with the bias+ReLU epilogue; the whole step "
                    "incl. the attention decoder and optimizer is timed"}


RUNNERS = {"transformer": run_transformer, "beam": run_beam, "captioning": run_captioning}

if __name__ == "__main__":
    for name in ([a for a in sys.argv[1:] if not a.startswith('--')] or list(RUNNERS)):
        t0 = time.perf_counter()
        result = RUNNERS[name](breakdown=True) if name == "transformer" and "--breakdown" in sys.argv else RUNNERS[name]()
        result["wall_s"] = time.perf_counter() - t0
        print(json.dumps(result), flush=True)

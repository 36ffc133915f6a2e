"""Independent second opinion on the oracle: HF transformers' port of the same published model
(feature extractor + WhisperModel with the oracle's weights remapped).  Not the reference, but different code."""
import numpy as np
import pytest
import torch

from b200_whisper.synth import MODEL_DIMS, random_state_dict, synth_audio
from oracle import whisper_oracle as wo
from oracle.tables import layout_for_vocab

transformers = pytest.importorskip("transformers")


@pytest.mark.parametrize("n_mels", [80, 128])
def test_mel_matches_hf_feature_extractor(n_mels):
    from transformers import WhisperFeatureExtractor
    from transformers.audio_utils import mel_filter_bank

    ours = wo.mel_filters(n_mels)
    hf = mel_filter_bank(201, n_mels, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
    assert np.abs(ours - hf).max() < 1e-7
    audio = synth_audio(4, 7.3)
    fe = WhisperFeatureExtractor(feature_size=n_mels)
    ref = np.asarray(fe._torch_extract_fbank_features(np.concatenate([audio, np.zeros(480000, np.float32)])[None]))[0]
    got = wo.log_mel_spectrogram(audio, n_mels, padding=480000).numpy()
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-5


def _to_hf(state, dims):
    from transformers import WhisperConfig, WhisperModel

    cfg = WhisperConfig(vocab_size=dims.n_vocab, num_mel_bins=dims.n_mels, d_model=dims.n_audio_state,
                        encoder_layers=dims.n_audio_layer, decoder_layers=dims.n_text_layer,
                        encoder_attention_heads=dims.n_audio_head, decoder_attention_heads=dims.n_text_head,
                        encoder_ffn_dim=4 * dims.n_audio_state, decoder_ffn_dim=4 * dims.n_text_state,
                        max_source_positions=1500, max_target_positions=448, activation_function="gelu")
    cfg._attn_implementation = "eager"
    m = WhisperModel(cfg).eval()
    sd = {}
    rn = {"attn.query": "self_attn.q_proj", "attn.key": "self_attn.k_proj", "attn.value": "self_attn.v_proj",
          "attn.out": "self_attn.out_proj", "attn_ln": "self_attn_layer_norm", "cross_attn.query": "encoder_attn.q_proj",
          "cross_attn.key": "encoder_attn.k_proj", "cross_attn.value": "encoder_attn.v_proj", "cross_attn.out": "encoder_attn.out_proj",
          "cross_attn_ln": "encoder_attn_layer_norm", "mlp.0": "fc1", "mlp.2": "fc2", "mlp_ln": "final_layer_norm"}
    for k, v in state.items():
        side, rest = k.split(".", 1)
        if rest.startswith("blocks."):
            _, i, tail = rest.split(".", 2)
            for a, b in sorted(rn.items(), key=lambda kv: -len(kv[0])):
                if tail.startswith(a + "."):
                    tail = b + tail[len(a):]
                    break
            sd[f"{side}.layers.{i}.{tail}"] = v
        elif rest.startswith("conv"):
            sd[f"{side}.{rest}"] = v
        elif rest == "ln_post.weight" or rest == "ln_post.bias" or rest.startswith("ln."):
            sd[f"{side}.layer_norm.{rest.split('.')[-1]}"] = v
        elif rest == "token_embedding.weight":
            sd["decoder.embed_tokens.weight"] = v
        elif rest == "positional_embedding":
            sd[f"{side}.embed_positions.weight"] = v
    sd["encoder.embed_positions.weight"] = wo.sinusoids(1500, dims.n_audio_state)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("k_proj.bias" in k for k in missing), (missing, unexpected)
    for k, p in m.named_parameters():  # HF key projections carry a bias openai's do not
        if "k_proj.bias" in k:
            torch.nn.init.zeros_(p)
    return m


def test_encoder_decoder_match_hf_whisper():
    dims = MODEL_DIMS["test-tiny"]
    state = random_state_dict(dims, 0, emb_std=0.1)
    model = wo.Whisper(wo.ModelDimensions(**dims.__dict__), state)
    hf = _to_hf(state, dims)
    audio = synth_audio(6, 5.0)
    mel = wo.pad_or_trim(wo.log_mel_spectrogram(audio, 80, padding=480000), 3000)[None]
    lay = layout_for_vocab(dims.n_vocab)
    tokens = torch.tensor([[lay.sot, lay.language_token("en"), lay.transcribe, lay.timestamp_begin, 100, 2000, 31000]])
    with torch.no_grad():
        xa = model.encode(mel)
        logits = model.decode(tokens, xa)
        enc_hf = hf.encoder(mel).last_hidden_state
        dec_hf = hf.decoder(input_ids=tokens, encoder_hidden_states=enc_hf).last_hidden_state
        logits_hf = dec_hf @ hf.decoder.embed_tokens.weight.t()
    assert (xa - enc_hf).abs().max() < 2e-4
    assert (logits - logits_hf).abs().max() < 2e-3
    assert logits.argmax(-1).tolist() == logits_hf.argmax(-1).tolist()


@pytest.mark.parametrize("n_vocab", [51864, 51865, 51866])
def test_logit_filters_match_hf_logits_processors(n_vocab):
    """SuppressBlank + SuppressTokens + ApplyTimestampRules of the oracle (`_Filters`, restating upstream decoding.py) against
    HF's ports of the same three filters (generation/logits_process.py) on random logits and token histories that reach every
    branch: first sampled position, text after a timestamp, a closed pair, an open timestamp, non-decreasing timestamps,
    timestamp mass above / below the best text token."""
    from types import SimpleNamespace

    from transformers.generation.logits_process import (SuppressTokensAtBeginLogitsProcessor, SuppressTokensLogitsProcessor,
                                                        WhisperTimeStampLogitsProcessor)

    lay = layout_for_vocab(n_vocab)
    tb, V = lay.timestamp_begin, lay.n_vocab
    initial = list(lay.sot_sequence("en", "transcribe"))
    begin = len(initial)
    cfg = SimpleNamespace(no_timestamps_token_id=lay.no_timestamps, eos_token_id=lay.eot, bos_token_id=lay.eot,
                          max_initial_timestamp_index=50, _detect_timestamp_from_logprob=True)
    hf = [SuppressTokensAtBeginLogitsProcessor([lay.blank, lay.eot], begin), SuppressTokensLogitsProcessor(list(lay.suppress_tokens())),
          WhisperTimeStampLogitsProcessor(cfg, begin_index=begin)]
    f = wo._Filters(lay, begin, wo.DecodingOptions(), n_audio_ctx=1500)
    assert f.max_initial_timestamp_index == 50
    rng = np.random.default_rng(n_vocab)
    text = lambda: int(rng.integers(0, lay.eot))
    histories = [[], [tb + 3], [tb + 3, text()], [tb + 3, text(), text()], [tb + 3, text(), tb + 40], [tb + 3, text(), tb + 40, tb + 40],
                 [tb, text(), tb + 10, tb + 10, text(), text(), tb + 700], [text()], [text(), text(), tb + 1499], [tb + 1500],
                 [tb + 5, tb + 5], [tb + 20, text(), tb + 20, tb + 20, text()]]
    for hist in histories:
        for trial in range(6):
            logits = torch.from_numpy(rng.normal(size=(2, V)).astype(np.float32)) * 3.0
            if trial % 3 == 1:
                logits[:, text()] += 40.0          # text mass wins the last rule
            if trial % 3 == 2:
                logits[:, tb + 800 : tb + 900] += 6.0  # timestamp mass wins it
            tokens = torch.tensor([initial + hist, initial + hist])
            want = logits.clone()
            for proc in hf:
                want = proc(tokens, want)
            got = logits.clone()
            f.apply(got, tokens)
            assert torch.equal(torch.isfinite(got), torch.isfinite(want)), (hist, trial)
            assert torch.equal(torch.where(torch.isfinite(got), got, torch.zeros(())), torch.where(torch.isfinite(want), want, torch.zeros(())))


def test_greedy_decode_matches_hf_model_with_hf_cache_and_processors():
    """The whole greedy loop in composition: HF's model code with HF's own KV cache and HF's logits processors, driven by a
    plain arg-max loop, must emit the oracle's token stream (`decode_window`: upstream DecodingTask + GreedyDecoder)."""
    from types import SimpleNamespace

    from transformers.generation.logits_process import (SuppressTokensAtBeginLogitsProcessor, SuppressTokensLogitsProcessor,
                                                        WhisperTimeStampLogitsProcessor)

    dims = MODEL_DIMS["test-tiny"]
    n_steps = 48
    for seed, eot_bias in ((6, 0.0), (8, 4.0)):
        state = random_state_dict(dims, 0, emb_std=0.1, eot_bias=eot_bias)
        model = wo.Whisper(wo.ModelDimensions(**dims.__dict__), state)
        hf = _to_hf(state, dims)
        lay = layout_for_vocab(dims.n_vocab)
        mel = wo.pad_or_trim(wo.log_mel_spectrogram(synth_audio(seed, 5.0), 80, padding=480000), 3000)
        want = wo.decode_window(model, mel, wo.DecodingOptions(language="en", sample_len=n_steps))
        assert want.min_margin > 1e-3
        initial = list(lay.sot_sequence("en", "transcribe"))
        cfg = SimpleNamespace(no_timestamps_token_id=lay.no_timestamps, eos_token_id=lay.eot, bos_token_id=lay.eot,
                              max_initial_timestamp_index=50, _detect_timestamp_from_logprob=True)
        procs = [SuppressTokensAtBeginLogitsProcessor([lay.blank, lay.eot], len(initial)), SuppressTokensLogitsProcessor(list(lay.suppress_tokens())),
                 WhisperTimeStampLogitsProcessor(cfg, begin_index=len(initial))]
        with torch.no_grad():
            enc = hf.encoder(mel[None]).last_hidden_state
            tokens = torch.tensor([initial])
            past = None
            sum_logprob = 0.0
            for i in range(n_steps):
                out = hf.decoder(input_ids=tokens if past is None else tokens[:, -1:], encoder_hidden_states=enc,
                                 past_key_values=past, use_cache=True)
                past = out.past_key_values
                logits = (out.last_hidden_state[:, -1] @ hf.decoder.embed_tokens.weight.t()).float()
                for proc in procs:
                    logits = proc(tokens, logits)
                nxt = int(logits.argmax(-1))
                sum_logprob += float(torch.log_softmax(logits, -1)[0, nxt])
                tokens = torch.cat([tokens, torch.tensor([[nxt]])], dim=-1)
                if nxt == lay.eot:
                    break
        got = tokens[0, len(initial):].tolist()
        if got and got[-1] == lay.eot:
            got = got[:-1]
        assert got == want.tokens, (seed, eot_bias)
        assert abs(sum_logprob - want.sum_logprob) < 2e-3 * max(1.0, abs(want.sum_logprob))

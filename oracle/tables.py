"""Token-id layout and suppress tables for the oracle (test infrastructure).

Restates openai-whisper 20250625 `tokenizer.py` special-token layout
(SURVEY.md Appendix A.6).  The vocab assets (`gpt2.tiktoken`,
`multilingual.tiktoken`) are absent, so the non-speech symbol ids are taken from
HF `transformers/models/whisper/configuration_whisper.py:23-44`, an independent
port of the same published tables, minus the special ids HF appended.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

LANGUAGES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no "
    "th ur hr bg lt la mi ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw "
    "gl mr pa si km sn yo so af oc ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl "
    "mg as tt haw ln ha ba jw su yue"
).split()
assert len(LANGUAGES) == 100


@dataclass(frozen=True)
class TokenLayout:
    n_vocab: int
    multilingual: bool
    num_languages: int
    eot: int
    sot: int
    translate: int
    transcribe: int
    sot_lm: int
    sot_prev: int
    no_speech: int
    no_timestamps: int
    timestamp_begin: int
    blank: int  # encode(" ")[0]
    non_speech: Tuple[int, ...]

    def language_token(self, code: str) -> int:
        idx = LANGUAGES.index(code)
        if idx >= self.num_languages:
            raise KeyError(code)
        return self.sot + 1 + idx

    @property
    def all_language_tokens(self) -> Tuple[int, ...]:
        return tuple(self.sot + 1 + i for i in range(self.num_languages))

    def sot_sequence(self, language: str | None, task: str | None) -> Tuple[int, ...]:
        # tokenizer.py: sot_sequence = [sot] (+ lang, + task when language/task given; only
        # multilingual tokenizers are given them)
        if not self.multilingual:
            return (self.sot,)
        seq = [self.sot, self.language_token(language or "en")]
        seq.append(self.transcribe if (task or "transcribe") == "transcribe" else self.translate)
        return tuple(seq)

    def suppress_tokens(self) -> Tuple[int, ...]:
        # decoding.py DecodingTask._get_suppress_tokens with suppress_tokens="-1"
        s = set(self.non_speech)
        s.update([self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm, self.no_speech])
        return tuple(sorted(s))


def layout_for_vocab(n_vocab: int) -> TokenLayout:
    from transformers.models.whisper.configuration_whisper import (
        NON_SPEECH_TOKENS,
        NON_SPEECH_TOKENS_MULTI,
    )

    multilingual = n_vocab >= 51865
    num_languages = n_vocab - 51765 - int(multilingual)
    eot = 50257 if multilingual else 50256
    sot = eot + 1
    translate = sot + 1 + num_languages
    non_speech = tuple(t for t in (NON_SPEECH_TOKENS_MULTI if multilingual else NON_SPEECH_TOKENS) if t < 50256)
    return TokenLayout(
        n_vocab=n_vocab,
        multilingual=multilingual,
        num_languages=num_languages,
        eot=eot,
        sot=sot,
        translate=translate,
        transcribe=translate + 1,
        sot_lm=translate + 2,
        sot_prev=translate + 3,
        no_speech=translate + 4,
        no_timestamps=translate + 5,
        timestamp_begin=translate + 6,
        blank=220,
        non_speech=non_speech,
    )

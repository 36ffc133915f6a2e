"""Token-id layout, suppress tables and (optional) detokeniser for the three Whisper vocabularies.

Follows upstream `whisper/tokenizer.py` (openai-whisper 20250625, the package behind reference
`stt_server/model/backends/torch_whisper.py:21`).  The rank files (`gpt2.tiktoken`,
`multilingual.tiktoken`) ship inside that package, not with the server; point
`B200_WHISPER_VOCAB_DIR` at a directory holding them to get real text.  Without them token ids are
rendered as `<id>` placeholders (segment timing, language and token streams are unaffected).
"""
from __future__ import annotations

import base64
import logging
import os
from dataclasses import dataclass
from functools import lru_cache
from typing import List, Optional, Sequence, Tuple

LOGGER = logging.getLogger("stt_server.model_backend")

LANGUAGES = {
    "en": "english", "zh": "chinese", "de": "german", "es": "spanish", "ru": "russian", "ko": "korean",
    "fr": "french", "ja": "japanese", "pt": "portuguese", "tr": "turkish", "pl": "polish", "ca": "catalan",
    "nl": "dutch", "ar": "arabic", "sv": "swedish", "it": "italian", "id": "indonesian", "hi": "hindi",
    "fi": "finnish", "vi": "vietnamese", "he": "hebrew", "uk": "ukrainian", "el": "greek", "ms": "malay",
    "cs": "czech", "ro": "romanian", "da": "danish", "hu": "hungarian", "ta": "tamil", "no": "norwegian",
    "th": "thai", "ur": "urdu", "hr": "croatian", "bg": "bulgarian", "lt": "lithuanian", "la": "latin",
    "mi": "maori", "ml": "malayalam", "cy": "welsh", "sk": "slovak", "te": "telugu", "fa": "persian",
    "lv": "latvian", "bn": "bengali", "sr": "serbian", "az": "azerbaijani", "sl": "slovenian", "kn": "kannada",
    "et": "estonian", "mk": "macedonian", "br": "breton", "eu": "basque", "is": "icelandic", "hy": "armenian",
    "ne": "nepali", "mn": "mongolian", "bs": "bosnian", "kk": "kazakh", "sq": "albanian", "sw": "swahili",
    "gl": "galician", "mr": "marathi", "pa": "punjabi", "si": "sinhala", "km": "khmer", "sn": "shona",
    "yo": "yoruba", "so": "somali", "af": "afrikaans", "oc": "occitan", "ka": "georgian", "be": "belarusian",
    "tg": "tajik", "sd": "sindhi", "gu": "gujarati", "am": "amharic", "yi": "yiddish", "lo": "lao",
    "uz": "uzbek", "fo": "faroese", "ht": "haitian creole", "ps": "pashto", "tk": "turkmen", "nn": "nynorsk",
    "mt": "maltese", "sa": "sanskrit", "lb": "luxembourgish", "my": "myanmar", "bo": "tibetan", "tl": "tagalog",
    "mg": "malagasy", "as": "assamese", "tt": "tatar", "haw": "hawaiian", "ln": "lingala", "ha": "hausa",
    "ba": "bashkir", "jw": "javanese", "su": "sundanese", "yue": "cantonese",
}
LANGUAGE_CODES = list(LANGUAGES.keys())
TO_LANGUAGE_CODE = {
    **{name: code for code, name in LANGUAGES.items()},
    "burmese": "my", "valencian": "ca", "flemish": "nl", "haitian": "ht", "letzeburgesch": "lb", "pushto": "ps",
    "panjabi": "pa", "moldavian": "ro", "moldovan": "ro", "sinhalese": "si", "castilian": "es", "mandarin": "zh",
}

# non-speech symbol ids of the two rank files (tokenizer.non_speech_tokens)
_NON_SPEECH_GPT2 = (
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 357, 366, 438, 532, 685,
    705, 796, 930, 1058, 1220, 1267, 1279, 1303, 1343, 1377, 1391, 1635, 1782, 1875, 2162, 2361, 2488, 3467, 4008,
    4211, 4600, 4808, 5299, 5855, 6329, 7203, 9609, 9959, 10563, 10786, 11420, 11709, 11907, 13163, 13697, 13700,
    14808, 15306, 16410, 16791, 17992, 19203, 19510, 20724, 22305, 22935, 27007, 30109, 30420, 33409, 34949, 40283,
    40493, 40549, 47282, 49146)
_NON_SPEECH_MULTI = (
    1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542, 873,
    893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585,
    6647, 7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553,
    16604, 18362, 18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470, 36865, 42863,
    47425, 49870, 50254)


@dataclass(frozen=True)
class Vocab:
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
    blank: Tuple[int, ...]
    non_speech: Tuple[int, ...]

    @property
    def first_language_token(self) -> int:
        return self.sot + 1

    def language_token(self, code: str) -> int:
        idx = LANGUAGE_CODES.index(code)
        if idx >= self.num_languages:
            raise ValueError(f"language {code} not in this vocabulary")
        return self.sot + 1 + idx

    def language_of_token(self, token: int) -> str:
        return LANGUAGE_CODES[token - self.sot - 1]

    def sot_sequence(self, language: Optional[str], task: Optional[str]) -> List[int]:
        if not self.multilingual:
            return [self.sot]
        return [self.sot, self.language_token(language or "en"),
                self.transcribe if (task or "transcribe") == "transcribe" else self.translate]

    def suppress_tokens(self) -> List[int]:
        """DecodingTask._get_suppress_tokens for the default suppress_tokens="-1"."""
        s = set(self.non_speech)
        s.update((self.transcribe, self.translate, self.sot, self.sot_prev, self.sot_lm, self.no_speech))
        return sorted(s)


@lru_cache(maxsize=None)
def vocab_for(n_vocab: int) -> Vocab:
    multilingual = n_vocab >= 51865
    num_languages = n_vocab - 51765 - int(multilingual)
    eot = 50257 if multilingual else 50256
    translate = eot + 2 + num_languages
    return Vocab(
        n_vocab=n_vocab, multilingual=multilingual, num_languages=num_languages, eot=eot, sot=eot + 1,
        translate=translate, transcribe=translate + 1, sot_lm=translate + 2, sot_prev=translate + 3,
        no_speech=translate + 4, no_timestamps=translate + 5, timestamp_begin=translate + 6,
        blank=(220,), non_speech=_NON_SPEECH_MULTI if multilingual else _NON_SPEECH_GPT2)


def normalize_language(language: Optional[str]) -> Optional[str]:
    """tokenizer.get_tokenizer's language handling: codes or english names, lower-cased."""
    if language is None:
        return None
    language = language.lower()
    if language not in LANGUAGES:
        if language in TO_LANGUAGE_CODE:
            return TO_LANGUAGE_CODE[language]
        raise ValueError(f"Unsupported language: {language}")
    return language


_WARNED: set = set()


def find_rank_file(name: str) -> Optional[str]:
    """`<name>.tiktoken`: under $B200_WHISPER_VOCAB_DIR, else in the assets of an installed openai-whisper (the package
    the reference's torch_whisper backend imports; located without importing it)."""
    roots = []
    if os.environ.get("B200_WHISPER_VOCAB_DIR"):
        roots.append(os.environ["B200_WHISPER_VOCAB_DIR"])
    try:
        import importlib.util

        spec = importlib.util.find_spec("whisper")
        if spec is not None and spec.submodule_search_locations:
            roots.extend(os.path.join(loc, "assets") for loc in spec.submodule_search_locations)
    except (ImportError, ValueError):
        pass
    for root in roots:
        path = os.path.join(root, f"{name}.tiktoken")
        if os.path.isfile(path):
            return path
    return None


class Detokenizer:
    """tiktoken-backed text codec.  Without the rank file: `<id>` placeholders if `allow_placeholders` (random-init models,
    tests, bench), else a RuntimeError -- a real checkpoint must never hand placeholder strings to clients."""

    def __init__(self, vocab: Vocab, allow_placeholders: bool = True):
        self.vocab = vocab
        self.encoding = None
        name = "multilingual" if vocab.multilingual else "gpt2"
        path = find_rank_file(name)
        if path is None and not allow_placeholders:
            raise RuntimeError(
                f"b200_whisper: tokenizer rank file {name}.tiktoken not found (set B200_WHISPER_VOCAB_DIR to the directory that "
                "holds it -- it ships in openai-whisper's whisper/assets/ -- or install openai-whisper next to this backend). "
                "Refusing to serve a real checkpoint with placeholder text; B200_WHISPER_ALLOW_PLACEHOLDER_TEXT=1 overrides.")
        if path:
            import tiktoken

            with open(path) as fh:
                ranks = {base64.b64decode(tok): int(rank) for tok, rank in (line.split() for line in fh if line.strip())}
            n = len(ranks)
            specials = ["<|endoftext|>", "<|startoftranscript|>",
                        *[f"<|{lang}|>" for lang in LANGUAGE_CODES[: vocab.num_languages]],
                        "<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>",
                        "<|notimestamps|>", *[f"<|{i * 0.02:.2f}|>" for i in range(1501)]]
            special_tokens = {}
            for tok in specials:
                special_tokens[tok] = n
                n += 1
            self.encoding = tiktoken.Encoding(
                name=os.path.basename(path), explicit_n_vocab=n,
                pat_str=r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
                mergeable_ranks=ranks, special_tokens=special_tokens)
        elif name not in _WARNED:
            _WARNED.add(name)
            LOGGER.warning("b200_whisper: no %s.tiktoken under B200_WHISPER_VOCAB_DIR; text is rendered as <id> placeholders", name)

    @property
    def has_text(self) -> bool:
        return self.encoding is not None

    def decode(self, tokens: Sequence[int]) -> str:
        toks = [t for t in tokens if t < self.vocab.timestamp_begin]
        if self.encoding is not None:
            return self.encoding.decode(toks)
        return "".join(f"<{t}>" for t in toks)

    def encode(self, text: str) -> Optional[List[int]]:
        if self.encoding is None:
            return None
        return self.encoding.encode(text)

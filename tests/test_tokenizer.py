"""The library's own SentencePiece-BPE tokenizer (kuiper/source/op/spm_bpe.cpp behind
op::SpeEncodeLayer, the front end of model::LLama2Model::encode / decode) against the SentencePiece
Python package: a Llama-style model (BPE, byte fallback, identity normalisation, dummy prefix) is
trained here on a small corpus, then ids and decoded text must match exactly."""
import random
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, str(ROOT / "kuiperllama_b200" / "kuiper"))
import build_host  # noqa: E402

spm = pytest.importorskip("sentencepiece")


def _corpus():
    import this  # noqa: F401  (the Zen of Python, rot13 in this.s)
    import codecs
    import inspect
    import json as _json
    import textwrap
    text = [codecs.decode(this.s, "rot13")]
    for mod in (textwrap, _json, inspect, random, subprocess):
        text.append(inspect.getsource(mod))
    extra = ["Once upon a time, there was a little girl named Lily. She loved to play outside.",
             "naïve café déjà vu — über straße", "東京は日本の首都です。", "Привет, мир! Как дела?",
             "1234567890 3.14159 2^10 = 1024", "tabs\tand  double  spaces   here"]
    return "\n".join(text).splitlines() + extra * 20


@pytest.fixture(scope="module")
def models(tmp_path_factory):
    d = tmp_path_factory.mktemp("spm")
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(l for l in _corpus() if l.strip()), encoding="utf-8")
    out = {}
    # (a) the Llama-2 recipe: BPE, byte fallback, identity normaliser, digits split, no squeezing
    spm.SentencePieceTrainer.train(
        input=str(corpus), model_prefix=str(d / "llama_like"), vocab_size=900, model_type="bpe",
        byte_fallback=True, normalization_rule_name="identity", add_dummy_prefix=True,
        remove_extra_whitespaces=False, split_digits=True, character_coverage=0.995,
        allow_whitespace_only_pieces=True, minloglevel=2)
    out["llama_like"] = d / "llama_like.model"
    # (b) no byte fallback, whitespace squeezing on: unknown characters become <unk>
    spm.SentencePieceTrainer.train(
        input=str(corpus), model_prefix=str(d / "plain"), vocab_size=500, model_type="bpe",
        byte_fallback=False, normalization_rule_name="identity", remove_extra_whitespaces=True,
        character_coverage=0.98, minloglevel=2)
    out["plain"] = d / "plain.model"
    # (c) what must be refused: a unigram model
    spm.SentencePieceTrainer.train(
        input=str(corpus), model_prefix=str(d / "unigram"), vocab_size=400, model_type="unigram",
        normalization_rule_name="identity", minloglevel=2)
    out["unigram"] = d / "unigram.model"
    return out


def _exe():
    exe = build_host.binary("llama2", "kuiper_tokenize")
    if not exe.exists():
        build_host.build("llama2")
    return str(exe)


def _sentences():
    rng = random.Random(7)
    base = [
        "Hello world", "hello", " leading space", "trailing space ", "two  spaces", "   ", "a", "",
        "Once upon a time, there was a little girl named Lily.", "The quick brown fox jumps over the lazy dog!",
        "naïve café — déjà vu", "東京は日本の首都です。", "Привет, мир!", "emoji 🙂 and 🚀 rockets",
        "mixed 東京 and English 123", "tab\there", "line one\\nline two", "x = y ** 2 + 3.14159;",
        "def f(a, b):\\n    return a + b", "ÿþý odd latin-1 letters", "▁ the blank symbol itself",
        "UPPER lower MiXeD", "a" * 50, "ab" * 40, "!!!???...", "https://example.com/path?q=1&r=2",
    ]
    words = ("the of and to in is it you that he was for on are with as his they be at one have this from or had by "
             "hot word but what some we can out other were all there when up use your how said an each she").split()
    for _ in range(60):
        base.append(" ".join(rng.choice(words) for _ in range(rng.randint(1, 14))))
    for _ in range(30):  # random unicode soup
        base.append("".join(chr(rng.choice([rng.randint(32, 126), rng.randint(0xA1, 0x17F), rng.randint(0x400, 0x44F),
                                                rng.randint(0x4E00, 0x4E80), 0x20])) for _ in range(rng.randint(1, 30))))
    return base


@pytest.mark.parametrize("name", ["llama_like", "plain"])
def test_encode_and_decode_match_sentencepiece(models, name):
    sp = spm.SentencePieceProcessor(model_file=str(models[name]))
    sents = [s for s in _sentences() if "\n" not in s]
    payload = "\n".join(sents) + "\n"
    r = subprocess.run([_exe(), str(models[name]), "encode"], input=payload, capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr
    got = [[int(t) for t in line.split()] for line in r.stdout.split("\n")[:len(sents)]]
    assert len(got) == len(sents)
    all_ids = []
    for s, ids in zip(sents, got):
        want = [sp.bos_id()] + sp.encode(s.replace("\\n", "\n"))
        assert ids == want, (name, s)
        all_ids.append(want)
    # decode: whole sentences, their BOS-less tails, and every single id of the vocabulary
    cases = all_ids + [ids[1:] for ids in all_ids] + [[i] for i in range(sp.get_piece_size())]
    rng = random.Random(11)  # arbitrary id sequences: byte pieces, control pieces and blanks in any order
    cases += [[rng.randrange(sp.get_piece_size()) for _ in range(rng.randint(1, 12))] for _ in range(400)]
    cases += [[rng.randrange(min(270, sp.get_piece_size())) for _ in range(rng.randint(1, 8))] for _ in range(300)]
    cases = [c for c in cases if c]
    r = subprocess.run([_exe(), str(models[name]), "decode"],
                       input="\n".join(" ".join(map(str, c)) for c in cases) + "\n", capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.split("\n")[:len(cases)]
    for c, line in zip(cases, lines):
        assert bytes.fromhex(line) == sp.decode(c).encode("utf-8"), (name, c)


def test_info_and_eos(models):
    sp = spm.SentencePieceProcessor(model_file=str(models["llama_like"]))
    r = subprocess.run([_exe(), str(models["llama_like"]), "info"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["vocab", str(sp.get_piece_size()), "eos_is_2", "1"]


def test_unsupported_or_missing_models_are_fatal(models, tmp_path):
    for path in (models["unigram"], tmp_path / "missing.model"):
        r = subprocess.run([_exe(), str(path), "info"], capture_output=True, text=True, timeout=60)
        assert r.returncode != 0
        assert "token model path is not valid" in r.stderr
    junk = tmp_path / "junk.model"
    junk.write_bytes(b"\x00\x01not a protobuf at all" * 10)
    r = subprocess.run([_exe(), str(junk), "info"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0


# ---- byte-level BPE (Llama-3 / Qwen2 tokenizer.json) against the Hugging Face `tokenizers` package ----------

LLAMA3_PAT = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|"
              r"\s*[\r\n]+|\s+(?!\S)|\s+")
QWEN_PAT = LLAMA3_PAT.replace(r"\p{N}{1,3}", r"\p{N}")
SPECIALS = {"llama3": ["<|begin_of_text|>", "<|end_of_text|>", "<|eot_id|>", "<|start_header_id|>"],
            "qwen": ["<|endoftext|>", "<|im_start|>", "<|im_end|>"]}


@pytest.fixture(scope="module")
def bpe_models(tmp_path_factory):
    tk = pytest.importorskip("tokenizers")
    from tokenizers import Regex, Tokenizer, decoders, models as tmodels, pre_tokenizers, trainers
    d = tmp_path_factory.mktemp("bpe")
    corpus = [l for l in _corpus() if l.strip()]
    out = {}
    for family, pat in (("llama3", LLAMA3_PAT), ("qwen", QWEN_PAT)):
        tok = Tokenizer(tmodels.BPE(ignore_merges=(family == "llama3")))
        tok.pre_tokenizer = pre_tokenizers.Sequence([
            pre_tokenizers.Split(Regex(pat), behavior="isolated", invert=False),
            pre_tokenizers.ByteLevel(add_prefix_space=False, trim_offsets=True, use_regex=False)])
        tok.decoder = decoders.ByteLevel()
        trainer = trainers.BpeTrainer(vocab_size=1200, special_tokens=SPECIALS[family], show_progress=False,
                                      initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
        tok.train_from_iterator(corpus, trainer)
        path = d / f"{family}.json"
        tok.save(str(path))
        out[family] = (path, tok)
    return out


def _bpe_sentences(family):
    s = [x for x in _sentences() if "\n" not in x]
    s += ["it's DON'T we'Re I'LL they'd you'VE I'm", "12345678 1 22 333 4444 5.5 3,141", "a  b   c    d",
          "  leading and trailing   ", "line\\n\\nbreaks \\n  indented\\n", "tabs\t\tand\tspaces \t mix",
          "x+=1; y=[1,2,3] # comment!!", "‘quotes’ “double” …ellipsis — dash", "no_space_before(paren)",
          "日本語のテキスト123と数字", "Ünïcödé ßtraße ŒUVRE", "emoji 🙂🙂 🚀", " nbsp em space　ideographic",
          SPECIALS[family][0] + "hello" + SPECIALS[family][1], "text " + SPECIALS[family][-1] + " more text",
          "<|not_a_special|> <|", "'", "''s", "'S", " '", "a'b'c", "ſ 'ſ"]
    # fuzz: every class the split pattern distinguishes, in random order ("\\n" stands for a newline)
    rng = random.Random(23)
    alphabet = (list("abcXYZéßжщ日本ǅ") + list("0123456789²½٣") + list(".,;:!?()[]{}<>|+-*/=_#@&%$^~`\"") +
                ["'", "'s", "'T", "'re", "'LL", " ", " ", "  ", "\t", "\r", "\\n", "\\n", "\u00a0", "\u3000", "\u2009",
                 "🙂", "\u0301"])
    for _ in range(400):
        s.append("".join(rng.choice(alphabet) for _ in range(rng.randint(1, 24))))
    return s


@pytest.mark.parametrize("family", ["llama3", "qwen"])
def test_byte_bpe_matches_hf_tokenizers(bpe_models, family):
    path, tok = bpe_models[family]
    sents = _bpe_sentences(family)
    r = subprocess.run([_exe(), f"--{family}", str(path), "encode"], input="\n".join(sents) + "\n",
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = [[int(t) for t in line.split()] for line in r.stdout.split("\n")[:len(sents)]]
    bos = [tok.token_to_id("<|begin_of_text|>")] if family == "llama3" else []  # Qwen2: no BOS (model.cpp:111)
    all_ids = []
    for s, ids in zip(sents, got):
        want = bos + tok.encode(s.replace("\\n", "\n"), add_special_tokens=False).ids
        assert ids == want, (family, s)
        all_ids.append(want)
    rng = random.Random(5)
    cases = [c for c in all_ids if c] + [[i] for i in range(tok.get_vocab_size())]
    cases += [[rng.randrange(tok.get_vocab_size()) for _ in range(rng.randint(1, 10))] for _ in range(300)]
    r = subprocess.run([_exe(), f"--{family}", str(path), "decode"],
                       input="\n".join(" ".join(map(str, c)) for c in cases) + "\n", capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stderr
    for c, line in zip(cases, r.stdout.split("\n")):
        # raw bytes of the pieces; the Python decoder additionally replaces malformed UTF-8
        want = tok.decode(c, skip_special_tokens=False)
        assert bytes.fromhex(line).decode("utf-8", errors="replace") == want, (family, c)


def test_byte_bpe_info_and_errors(bpe_models, tmp_path):
    path, tok = bpe_models["qwen"]
    r = subprocess.run([_exe(), "--qwen", str(path), "info"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split()[:2] == ["vocab", str(tok.get_vocab_size())]
    bad = tmp_path / "bad.json"
    bad.write_text('{"model": {"type": "Unigram"}}')
    for p in (bad, tmp_path / "missing.json"):
        r = subprocess.run([_exe(), "--llama3", str(p), "info"], capture_output=True, text=True, timeout=60)
        assert r.returncode != 0 and "token model path is not valid" in r.stderr


@pytest.mark.parametrize("flag,vocab,bos", [(None, 32000, True), ("--llama3", 128256, True), ("--qwen", 151936, False)])
def test_stand_in_tokenizer_for_synthetic_checkpoints(flag, vocab, bos):
    """Path "<none>": ids in, "<id>" text out, never a sentence end -- what kuiper_decode and the GPU
    tests of the C++ model run with (synthetic checkpoints have no tokenizer file)."""
    base = [_exe()] + ([flag] if flag else []) + ["<none>"]
    r = subprocess.run(base + ["info"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.split() == ["vocab", str(vocab), "eos_is_2", "0"]
    r = subprocess.run(base + ["encode"], input="hi there\n", capture_output=True, text=True, timeout=60)
    ids = [int(t) for t in r.stdout.split()]
    assert (ids[0] == 1) == bos and len(ids) == len("hi there") + (1 if bos else 0)
    r = subprocess.run(base + ["decode"], input="5 6 7\n", capture_output=True, text=True, timeout=60)
    assert bytes.fromhex(r.stdout.strip()).decode() == "<5><6><7>"

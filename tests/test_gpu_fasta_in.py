"""Device-side FASTA reader (rlsim_add_fasta) against oracle/fasta_reader.py, the byte-by-byte
restatement of fasta.go:113-212 / input.go:82-124 that is pinned against the release binary (tests/test_oracle_pin.py):
same records, same skips, same sequences, same library."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util
from oracle import fasta_reader

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(util.HERE)


def _host_index(data, expr_mul=1.0):
    return [(r["line"], r["name"], r["level"], len(r["seq"]), r["seq"]) for r in fasta_reader.read(data, expr_mul)]


def _device(gpu, data, expr_mul=1.0, argv=("-n", "300", "-si", "2")):
    args = util.parse(list(argv))
    h = gpu.Handle(0)
    h.set_model(args.to_params())
    h.sample_target(args.req_frags)
    h.set_first_index(0)
    idx = h.add_fasta(data, expr_mul)
    return h, idx, args


def _check_index(gpu, data, expr_mul=1.0):
    rows = _host_index(data, expr_mul)
    h, idx, args = _device(gpu, data, expr_mul)
    assert len(idx["valid"]) == len(rows)
    for r, (line, name, level, slen, seq) in enumerate(rows):
        s0 = int(idx["name_start"][r])
        assert data[s0:s0 + int(idx["line_len"][r])] == line, r
        assert int(idx["seq_len"][r]) == slen, r
        assert bool(idx["valid"][r]) == (name is not None), (r, line)
        if name is not None:
            assert data[s0:s0 + int(idx["name_len"][r])] == name
            assert int(idx["level"][r]) == level
    assert idx["n_added"] == sum(1 for row in rows if row[1] is not None)
    # the library built from the device-parsed text equals the one built from the host-parsed records
    seqs = [row[4] for row in rows if row[1] is not None]
    levels = [row[2] for row in rows if row[1] is not None]
    names = [row[1] for row in rows if row[1] is not None]
    if not seqs:
        return
    h.build_library()
    h.sample()
    g = gpu.Handle(0)
    g.set_model(args.to_params())
    g.sample_target(args.req_frags)
    g.set_first_index(0)
    g.add_transcripts(seqs, levels)
    g.build_library()
    g.sample()
    for k in ("tr", "start", "end", "count0", "count"):
        assert np.array_equal(h.species()[k], g.species()[k]), k
    assert h.fasta(names) == g.fasta(names)


def _seq(rng, n, alphabet=b"ACGT"):
    return bytes(rng.choice(list(alphabet), size=n).astype(np.uint8))


def test_texts_pinned_against_the_release_binary(gpu):
    """The unusual texts of tests/golden/fasta_reader.json (what the binary made of them pins the oracle reader)."""
    import json
    for name, g in sorted(json.load(open(os.path.join(util.GOLDEN, "fasta_reader.json"))).items()):
        mul = float(g["argv"][g["argv"].index("-m") + 1]) if "-m" in g["argv"] else 1.0
        _check_index(gpu, g["text"].encode("latin1"), expr_mul=mul)


def test_basic_fixture_and_layouts(gpu):
    data = open(util.BASIC_FASTA, "rb").read()
    _check_index(gpu, data)
    _check_index(gpu, data, expr_mul=0.37)
    _check_index(gpu, data.replace(b"\n", b"\r\n"))
    _check_index(gpu, data.lower().replace(b">", b">"))  # lower-case names and sequences (names are case-preserved)


def test_reader_edge_cases(gpu):
    rng = np.random.default_rng(7)
    s = [_seq(rng, n) for n in (700, 1300, 5000, 17, 900, 4096 * 3 + 5, 800, 600)]

    def wrap(x, w):
        return b"\n".join(x[i:i + w] for i in range(0, len(x), w))

    text = b"".join([
        b">a$120\n" + wrap(s[0], 60) + b"\n",
        b">b two words$ 45 trailing junk\n" + wrap(s[1].lower(), 71) + b"\n\n\n",          # TrimSpace, Sscanf stops at junk
        b">no_dollar\n" + wrap(s[2], 80) + b"\n",                                           # skipped: malformed name
        b">x$1$2\n" + s[3] + b"\n",                                                         # skipped: two '$'
        b">$7\n" + s[3] + b"\n",                                                            # skipped: empty name
        b">neg$-5\n" + s[3] + b"\n",                                                        # skipped: Sscanf into uint64
        b">huge$99999999999999999999999\n" + s[3] + b"\n",                                  # skipped: out of range
        b">c$33\n" + wrap(s[4], 50).replace(b"\n", b" \t\r\n") + b"\n",                      # spaces / tabs / CR inside the sequence
        b">d$250\n" + wrap(s[5], 61) + b"\n",                                               # spans several 4 KiB tiles
        b">empty$9\n",                                                                      # no sequence
        b">e$77\n" + s[6][:300] + b">f$88\n" + s[7] + b"\n",                                # '>' in the middle of a line opens a record
        b">>g$5\n" + s[0][:400] + b"\n",                                                    # '>>': the second '>' is part of the name
        b">h$12\n" + s[1][:777],                                                            # last record without a final newline
    ])
    _check_index(gpu, text)
    _check_index(gpu, text + b"\n>tail$3")   # last record is a bare name line
    for cut in (1, 15, 16, 17, 4095, 4096, 4097):
        _check_index(gpu, text[:len(text) - cut])


def test_random_texts(gpu):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTacgtNn>$\n\n \t\r0123456789xyz", dtype=np.uint8)
    weights = np.array([8] * 8 + [1, 1, 0.15, 0.15, 1, 1, 0.1, 0.1, 0.1] + [0.3] * 10 + [0.2] * 3)
    weights = weights / weights.sum()
    for trial in range(12):
        n = int(rng.integers(1, 30000))
        body = rng.choice(alphabet, size=n, p=weights).tobytes()
        text = b">r$%d\n" % rng.integers(0, 50) + body
        rows = _host_index(text)
        h, idx, _ = _device(gpu, text)
        assert len(idx["valid"]) == len(rows), trial
        for r, (line, name, level, slen, seq) in enumerate(rows):
            s0 = int(idx["name_start"][r])
            assert text[s0:s0 + int(idx["line_len"][r])] == line, (trial, r)
            assert int(idx["seq_len"][r]) == slen, (trial, r)
            assert bool(idx["valid"][r]) == (name is not None), (trial, r, line)
            if name is not None:
                assert int(idx["level"][r]) == level


def test_illegal_text_and_long_runs(gpu):
    args = util.parse(["-n", "10"])
    h = gpu.Handle(0)
    h.set_model(args.to_params())
    with pytest.raises(gpu.RlsimError) as e:
        h.add_fasta(b"ACGT\n>a$1\nACGT\n")
    assert "Illegal fasta record" in str(e.value)   # fasta.go:183-185
    with pytest.raises(gpu.RlsimError) as e:
        h.add_fasta(b">a$1\nACGT\n" + b">" * 200 + b"b$2\nACGT\n")
    assert e.value.code == gpu.E_UNSUPPORTED         # the host reader takes such texts (main.py / main.cpp fall back)
    idx = h.add_fasta(b"")
    assert len(idx["valid"]) == 0


def test_cli_device_reader_equals_host_reader(tmp_path):
    def run(env_extra, exe):
        rep = str(tmp_path / ("rep_%s_%s.json" % (exe[-1][-6:], len(env_extra))))
        env = dict(os.environ, **env_extra)
        p = subprocess.run(exe + ["-r", rep, "-n", "5000", "-si", "9", "-c", "9", util.BASIC_FASTA], cwd=ROOT, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        assert p.returncode == 0, p.stderr.decode()[-2000:]
        return p.stdout, open(rep).read()

    py = [sys.executable, "-m", "rlsim_b200"]
    cc = [os.path.join(ROOT, "rlsim_b200", "rlsim_gpu")]
    a = run({}, py)
    b = run({"RLSIM_HOST_FASTA": "1"}, py)
    c = run({}, cc)
    d = run({"RLSIM_HOST_FASTA": "1"}, cc)
    assert a[0] == b[0] == c[0] == d[0] and len(a[0]) > 100000
    assert a[1] == b[1] and c[1] == d[1]

"""Host-side ingest of the C++ mirror (include/frecsys/dataset.h): the mmap + multi-threaded
parser must give the reference's tuple list — same order, same atoi/substr semantics on odd lines — and the
lazily built by_user() lists.  Checker: the oracle's Dataset::FromCsv (restates dataset.h:71-99) and a
line-by-line Python restatement.  CPU only."""
import os
import subprocess

import numpy as np
import pytest

import helpers


def _build_tool():
    subprocess.run(["make", "-C", os.path.join(helpers.ROOT, "tools"), "dataset_dump"], check=True,
                   capture_output=True)
    return os.path.join(helpers.ROOT, "tools", "dataset_dump")


def _load(tool, path, tmp_path, maps=False):
    out = os.path.join(tmp_path, "dump.bin")
    subprocess.run([tool, path, out] + (["maps"] if maps else []), check=True, capture_output=True)
    raw = np.fromfile(out, dtype=np.int32)
    n, mu, mi = raw[:3]
    users, items = raw[3:3 + n], raw[3 + n:3 + 2 * n]
    rest = raw[3 + 2 * n:]
    rows = {}
    k = 0
    while k < len(rest):
        r, m = rest[k], rest[k + 1]
        rows[int(r)] = rest[k + 2:k + 2 + 2 * m].reshape(m, 2)
        k += 2 + 2 * m
    return int(n), int(mu), int(mi), users, items, rows


def _atoi(s):
    s = s.lstrip(" \t\n\v\f\r")
    sign = 1
    if s[:1] in ("+", "-"):
        sign = -1 if s[0] == "-" else 1
        s = s[1:]
    d = ""
    for ch in s:
        if not ch.isdigit():
            break
        d += ch
    return sign * int(d) if d else 0


def _reference_parse(text):
    """getline loop of dataset.h:76-92 on the file content."""
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines.pop()          # getline does not produce an empty last line after the final '\n'
    out = []
    for line in lines[1:]:   # header dropped
        pos = line.find(",")
        u = _atoi(line[:pos] if pos >= 0 else line)
        i = _atoi(line[pos + 1:])   # pos == -1 -> substr(0): the whole line
        out.append((u, i))
    return out


def test_fixture_tuples_match_the_oracle(tmp_path):
    from oracle import loader as O
    tool = _build_tool()
    path = helpers.fixture_csv("train")
    n, mu, mi, users, items, rows = _load(tool, path, str(tmp_path), maps=True)
    ods = O.Dataset.from_csv(path)
    ou, oi = ods.tuples()
    assert n == len(ou) and np.array_equal(users, ou) and np.array_equal(items, oi)
    assert (mu, mi) == (ods.max_user, ods.max_item)
    ptr, ids, tup = ods.csr(False, mu + 1)
    for r in (0, 1, 17, mu):
        lo, hi = ptr[r], ptr[r + 1]
        if hi > lo:
            assert np.array_equal(rows[r][:, 0], ids[lo:hi]) and np.array_equal(rows[r][:, 1], tup[lo:hi])
        else:
            assert r not in rows
    assert len(rows) == ods.distinct_users


@pytest.mark.parametrize("body", [
    "1,2\n3,4\n",                       # plain
    "1,2\n3,4",                         # no trailing newline
    "1,2\r\n3,4\r\n",                   # CRLF
    " 7, 8\n\n9,10\n",                  # leading blanks, an empty line (-> tuple (0,0))
    "5\n6,7,8\n-3,+4\nx,y\n",           # no comma, extra field, signs, non-numeric
    "",                                 # header only
])
def test_odd_lines_follow_the_reference_semantics(tmp_path, body):
    tool = _build_tool()
    text = "uid,sid\n" + body
    path = os.path.join(str(tmp_path), "odd.csv")
    with open(path, "w", newline="") as f:
        f.write(text)
    n, mu, mi, users, items, _ = _load(tool, path, str(tmp_path))
    want = _reference_parse(text)
    assert n == len(want)
    assert [(int(a), int(b)) for a, b in zip(users, items)] == want


def test_large_file_is_parsed_in_parallel_in_file_order(tmp_path):
    """> 1 MiB so that the multi-threaded path (chunks cut at line starts) is taken."""
    tool = _build_tool()
    rng = np.random.default_rng(4)
    u = rng.integers(0, 200000, 300000)
    i = rng.integers(0, 50000, 300000)
    path = os.path.join(str(tmp_path), "big.csv")
    with open(path, "w") as f:
        f.write("uid,sid\n")
        f.write("\n".join(f"{a},{b}" for a, b in zip(u, i)))
        f.write("\n")
    assert os.path.getsize(path) > (1 << 20)
    n, mu, mi, users, items, _ = _load(tool, path, str(tmp_path))
    assert n == 300000 and np.array_equal(users, u) and np.array_equal(items, i)
    assert (mu, mi) == (int(u.max()), int(i.max()))


def test_binary_cache_round_trip(tmp_path):
    """FRECSYS_DATASET_CACHE=1 (binary cache of the parsed CSR input): the second load comes from
    `<csv>.frxbin` and is identical; a changed CSV (size / mtime key) invalidates the cache; a damaged cache is
    ignored; without the switch nothing is written next to the input."""
    tool = _build_tool()
    rng = np.random.default_rng(3)
    path = os.path.join(tmp_path, "d.csv")
    rows = [(int(u), int(i)) for u, i in zip(rng.integers(0, 300, 5000), rng.integers(0, 200, 5000))]
    with open(path, "w") as f:
        f.write("uid,sid\n" + "".join(f"{u},{i}\n" for u, i in rows))
    n0, _, _, u0, i0, _ = _load(tool, path, tmp_path)
    assert not os.path.exists(path + ".frxbin")
    env = dict(os.environ, FRECSYS_DATASET_CACHE="1")

    def load_cached():
        out = os.path.join(tmp_path, "dump.bin")
        subprocess.run([tool, path, out], check=True, capture_output=True, env=env)
        raw = np.fromfile(out, dtype=np.int32)
        n = int(raw[0])
        return n, raw[3:3 + n].copy(), raw[3 + n:3 + 2 * n].copy()

    n1, u1, i1 = load_cached()
    assert os.path.exists(path + ".frxbin")
    n2, u2, i2 = load_cached()          # from the cache
    assert n0 == n1 == n2 == 5000
    assert np.array_equal(u0, u1) and np.array_equal(u1, u2) and np.array_equal(i0, i1) and np.array_equal(i1, i2)
    # the cache really is what gets read: poison its payload (same header) and load again
    blob = bytearray(open(path + ".frxbin", "rb").read())
    blob[32:36] = (12345).to_bytes(4, "little")
    st = os.stat(path)
    open(path + ".frxbin", "wb").write(bytes(blob))
    os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns))
    n3, u3, _ = load_cached()
    assert n3 == 5000 and u3[0] == 12345
    # a changed CSV invalidates it
    with open(path, "a") as f:
        f.write("7,9\n")
    n4, u4, i4 = load_cached()
    assert n4 == 5001 and u4[-1] == 7 and i4[-1] == 9 and u4[0] == rows[0][0]
    # a truncated cache is ignored and rewritten
    open(path + ".frxbin", "wb").write(open(path + ".frxbin", "rb").read()[:100])
    n5, u5, _ = load_cached()
    assert n5 == 5001 and np.array_equal(u5, u4)
    n6, u6, _ = load_cached()
    assert n6 == 5001 and np.array_equal(u6, u4)

"""Native batch TextGrid reader (host threads) against the Python reader on generated files: long and short formats,
UTF-16, escaped quotes, zero-length intervals, a point tier, missing / broken files."""
import numpy as np


def _long(tiers, xmax):
    out = ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", "xmin = 0 ", f"xmax = {xmax} ", "tiers? <exists> ", f"size = {len(tiers)} ", "item []: "]
    for k, (name, ivs) in enumerate(tiers, 1):
        out += [f"    item [{k}]:", '        class = "IntervalTier" ', f'        name = "{name}" ', "        xmin = 0 ", f"        xmax = {xmax} ",
                f"        intervals: size = {len(ivs)} "]
        for j, (a, b, m) in enumerate(ivs, 1):
            m = m.replace('"', '""')
            out += [f"        intervals [{j}]:", f"            xmin = {a} ", f"            xmax = {b} ", f'            text = "{m}" ']
    return "\n".join(out) + "\n"


def _short(tiers, xmax):
    out = ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", "0", str(xmax), "<exists>", str(len(tiers))]
    for name, ivs in tiers:
        out += ['"IntervalTier"', f'"{name}"', "0", str(xmax), str(len(ivs))]
        for a, b, m in ivs:
            out += [str(a), str(b), '"' + m.replace('"', '""') + '"']
    return "\n".join(out) + "\n"


def test_native_batch_matches_python_reader(tmp_path):
    import prosody_b200  # noqa: F401
    from prosody_b200 import textgrid as TG
    rng = np.random.default_rng(3)
    words = ["bonjour", "à", "l'été", 'dit "non"', "", "[*]", "fin.", "où ?", "1,5", "x=3"]
    paths = []
    for f in range(150):
        t, ivs = 0.0, []
        for _ in range(int(rng.integers(1, 40))):
            d = float(rng.choice([0.0, 0.01, 0.123456789, 0.5, 1.25]))          # 0.0 -> dropped interval; 9 decimals -> rounding
            ivs.append((round(t, 9), round(t + d, 9), str(rng.choice(words)))); t += d
        tiers = [("words", ivs), ("phones", [(0, max(t, 0.01), "sil")])]
        text = (_long if f % 2 == 0 else _short)(tiers, max(t, 0.01))
        p = tmp_path / f"seg_{f}.TextGrid"
        if f % 7 == 3:
            p.write_bytes(b"\xff\xfe" + text.encode("utf-16-le"))
        elif f % 7 == 5:
            p.write_bytes(b"\xef\xbb\xbf" + text.encode("utf-8"))
        else:
            p.write_text(text, encoding="utf-8")
        paths.append(p)
    (tmp_path / "broken.TextGrid").write_text('File type = "ooTextFile"\nObject class = "TextGrid"\n0\n1\n<exists>\n1\n"IntervalTier"\n"words"\n0\n1\n3\n0\n0.5\n"a"\n', encoding="utf-8")
    (tmp_path / "other.txt").write_text("hello", encoding="utf-8")
    (tmp_path / "notier.TextGrid").write_text('File type = "ooTextFile"\nObject class = "TextGrid"\n0\n1\n<absent>\n', encoding="utf-8")
    paths += [tmp_path / "broken.TextGrid", tmp_path / "other.txt", tmp_path / "notier.TextGrid", tmp_path / "missing.TextGrid"]
    for tier in (0, 1):
        st, got = TG.read_tier_batch(paths, tier=tier, threads=4)
        first = TG.STATUS_NOT_TEXTGRID if tier == 0 else TG.STATUS_NO_TIER        # the truncated file has a single tier
        assert list(st[-4:]) == [first, TG.STATUS_NOT_TEXTGRID, TG.STATUS_NO_TIER, TG.STATUS_UNREADABLE]
        for p, s, ivs in zip(paths[:-4], st, got):
            assert s == TG.STATUS_OK
            assert ivs == list(TG.read(p).tiers[tier].intervals), p
    st1, got1 = TG.read_tier_batch(paths, tier=0, threads=1)
    assert np.array_equal(st1, TG.read_tier_batch(paths, tier=0, threads=4)[0]) and got1 == TG.read_tier_batch(paths, tier=0, threads=0)[1]
    st, got = TG.read_tier_batch([], tier=0)
    assert len(st) == 0 and got == []


def test_whisper_json_rules(tmp_path):
    """json_to_textgrid's rules (use_whisper_timestamped.py:330-395) and a round trip through writer and both readers."""
    import prosody_b200  # noqa: F401
    from prosody_b200 import textgrid as TG
    data = {"segments": [{"start": 0.2, "end": 1.4, "words": [{"text": "Bonjour", "start": 0.2, "end": 0.61},
                                                              {"text": "à", "start": 0.61, "end": 0.61},       # start >= end -> +0.01
                                                              {"text": "[*]", "start": 0.9, "end": 1.1}]},
                         {"start": 2.0, "end": 2.5, "words": [{"text": 'dit "oui"', "start": 2.0, "end": 2.5}]}]}
    ivs, xmax = TG.whisper_json_to_intervals(data)
    assert ivs == [(0.0, 0.2, " "), (0.2, 0.61, "Bonjour"), (0.61, 0.62, "à"), (0.62, 0.9, " "), (0.9, 1.1, " "), (1.1, 2.0, " "),
                   (2.0, 2.5, 'dit "oui"')] and xmax == 2.5
    assert TG.whisper_json_to_intervals({"segments": [{"start": 0.0, "end": 3.2, "words": []}]}) == ([(0.0, 3.2, "...")], 3.2)
    assert TG.whisper_json_to_intervals({"segments": []}) == ([(0.0, 1.0, "...")], 1.0)
    p = tmp_path / "w.TextGrid"
    TG.write_whisper_textgrid(p, data)
    back = TG.read(p)
    assert back.tiers[0].name == "words" and list(back.tiers[0].intervals) == ivs and back.xmax == 2.5
    st, batch = TG.read_tier_batch([p])
    assert st[0] == 0 and batch[0] == ivs

"""Row f4 (host text front end): throughput of chatttsplus_b200/text.py next to the reference's own commons/text_utils.py functions
on the same inputs, on this host's CPU (single thread).  The reference half needs /root/reference and is skipped when it is absent.
    python tests/prof_text.py"""
import importlib.util
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from chatttsplus_b200 import text as T  # noqa: E402

SENT = ["The quick brown fox jumps over 13 lazy dogs in 2024, and 105 more watched (quietly) from the hill.",
        "今天的气温是23度，明天会下雨；请带伞。我们在 3 点钟见面，好吗？[laugh] 真的很有趣……",
        "A considerably longer sentence, with several clauses; some of them short, others — like this one — rather winding, so that the splitter has "
        "something to do when the minimum length of 150 characters is reached and it has to look for the next punctuation mark. Then a short one."]


def bench(name, fn, args, n):
    t0 = time.perf_counter()
    for _ in range(n):
        for a in args:
            fn(a)
    dt = time.perf_counter() - t0
    print(f"  {name:28s} {n * len(args) / dt:12.0f} calls/s")
    return n * len(args) / dt


def run(mod, label, n=2000):
    print(label)
    out = {}
    out["num_to_english"] = bench("num_to_english", mod.num_to_english, [105, 2024, 13, 999999], n)
    out["num2text"] = bench("num2text", mod.num2text, SENT, n)
    out["remove_brackets"] = bench("remove_brackets", mod.remove_brackets, SENT, n)
    out["get_lang"] = bench("get_lang", mod.get_lang, SENT, n)
    out["split_text_by_punctuation"] = bench("split_text_by_punctuation", mod.split_text_by_punctuation, SENT, n)
    return out


ours = run(T, "chatttsplus_b200.text")
ref_path = "/root/reference/chattts_plus/commons/text_utils.py"
if os.path.exists(ref_path):
    import types
    zh = types.ModuleType("zh_normalization")   # absent third-party package, needed at import time only (as in tests/golden/make_golden.py)
    zh.TextNormalizer = object
    sys.modules.setdefault("zh_normalization", zh)
    spec = importlib.util.spec_from_file_location("ref_text_utils", ref_path)
    ref = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(ref)
        theirs = run(ref, "reference commons/text_utils.py")
        print("ratio ours / reference: " + ", ".join(f"{k} {ours[k] / theirs[k]:.2f}x" for k in ours))
    except Exception as e:   # e.g. a dependency of the reference module that is not installed here
        print("reference module not importable here:", type(e).__name__, e)

"""Host-side tokenizer of the keyword leg.

Behavioural twin of tokenize_french / FRENCH_STOPWORDS in the reference
(src/rag/bm25_index.py:22-49): lowercase, keep runs of [a-z0-9 + French
accented letters] joined by single hyphens, drop stop-words and tokens of
length <= 1.  tests/test_host.py pins it against tokens produced by the
reference itself (tests/golden/tokenizer.json).
"""
import re

_LETTERS = "a-zàâäéèêëïîôùûüÿçœæ0-9"
_TOKEN_RE = re.compile(f"[{_LETTERS}]+(?:-[{_LETTERS}]+)*")

STOPWORDS = frozenset("""
le la les de des du un une et en au aux ce ces cette qui que quoi dont où par pour dans sur avec sans sous
entre vers chez est sont être avoir fait faire peut il elle ils elles nous vous on se ne pas plus très aussi
mais ou donc car si ni je tu son sa ses leur leurs mon ma mes ton ta tes notre votre tout tous toute toutes
même autre autres quel quelle quels quelles comme été ayant après avant lors depuis pendant alors ainsi bien
peu trop assez encore déjà jamais rien chaque cet à d l n s c j m t y
""".split())


def tokenize_french(text):
    return [t for t in _TOKEN_RE.findall(text.lower()) if len(t) > 1 and t not in STOPWORDS]

"""Shared input generators for the parity tests (seeded, small enough for the CPU oracle)."""
import random

import numpy as np

from catch_b200 import genome, probe


def mutate(rng, s, rate, alphabet='ACGT'):
    return ''.join((rng.choice(alphabet) if rng.random() < rate else c) for c in s)


def random_groups(rng, alphabet='ACGT', n_groups=None, anc_len=(150, 500), max_genomes=6, with_n=True):
    """List (groups) of lists (genomes) of lists (sequence strings): diverged copies of one
    ancestor per group, some multi-sequence genomes, some Ns, some truncated sequences."""
    if n_groups is None:
        n_groups = rng.choice([1, 1, 2])
    groups = []
    for _ in range(n_groups):
        anc = ''.join(rng.choice(alphabet) for _ in range(rng.randint(*anc_len)))
        gens = []
        for _ in range(rng.randint(1, max_genomes)):
            seqs = []
            for _ in range(rng.choice([1, 1, 2, 3])):
                s = mutate(rng, anc, rng.choice([0.0, 0.02, 0.08]), alphabet)
                if with_n and rng.random() < 0.3:
                    s = list(s)
                    for _ in range(rng.randint(1, 4)):
                        s[rng.randrange(len(s))] = 'N'
                    s = ''.join(s)
                if rng.random() < 0.2:
                    s = s[:rng.randint(30, len(s))]
                seqs.append(s)
            gens.append(seqs)
        groups.append(gens)
    return groups


def to_genomes(groups):
    from collections import OrderedDict
    out = []
    for gens in groups:
        gl = []
        for seqs in gens:
            if len(seqs) == 1:
                gl.append(genome.Genome.from_one_seq(seqs[0]))
            else:
                gl.append(genome.Genome.from_chrs(OrderedDict((str(i), s) for i, s in enumerate(seqs))))
        out.append(gl)
    return out


def tile_candidates(seqs, probe_length, probe_stride):
    """Candidate probe strings from sequences (reference: filter/candidate_probes.py:97-106,
    without the N-string handling, which the callers here do not need)."""
    out = []
    for s in seqs:
        if len(s) < probe_length:
            out.append(s)
            continue
        for start in range(0, len(s), probe_stride):
            if start + probe_length > len(s):
                break
            out.append(s[start:start + probe_length])
        if len(s) % probe_stride != 0:
            out.append(s[len(s) - probe_length:])
    return out


def random_case(case, alphabet='ACGT'):
    """A full SetCoverFilter parity case: (groups, candidate strings per group, params)."""
    rng = random.Random(case)
    groups = random_groups(rng, alphabet=alphabet)
    pl = rng.choice([20, 30, 40])
    ps = rng.choice([5, 10, 15])
    params = dict(
        mismatches=rng.choice([0, 1, 2, 3]),
        lcf_thres=rng.choice([pl, pl, pl - 5, pl // 2]),
        island_of_exact_match=rng.choice([0, 0, 8]),
        cover_extension=rng.choice([0, 0, 10, 25]),
        coverage=rng.choice([1.0, 1.0, 0.8, 0.5, 120]),
        kmer_probe_map_k=rng.choice([10, 12, 20]),
    )
    cands = []
    for gens in groups:
        c = []
        for seqs in gens:
            c += [x for x in tile_candidates(seqs, pl, ps) if len(x) >= params['kmer_probe_map_k']]
        if rng.random() < 0.5:
            c = list(dict.fromkeys(c))
        cands.append(c)
    return groups, cands, params


def synthetic_genomes(n_genomes, length, div, seed):
    """SURVEY.md section 8(d) generator: one ancestor, i.i.d. substitutions to a different base."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, length)
    out = []
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)
    for _ in range(n_genomes):
        mask = rng.random(length) < div
        shift = rng.integers(1, 4, length)
        g = np.where(mask, (anc + shift) % 4, anc)
        out.append(letters[g].tobytes().decode())
    return out


INFLUENZA_SEGMENTS = [2341, 2341, 2233, 1778, 1565, 1413, 1027, 890]


def synthetic_influenza(n_genomes, seed, n_clades=10, clade_div=0.08, strain_div=0.02):
    """segmented genomes, one ancestor per segment, two-level
    divergence (clades, then strains).  Returns a list of genomes, each a list of 8 segment strings."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)

    def diverge(anc, div):
        mask = rng.random(len(anc)) < div
        shift = rng.integers(1, 4, len(anc))
        return np.where(mask, (anc + shift) % 4, anc)

    ancestors = [rng.integers(0, 4, L) for L in INFLUENZA_SEGMENTS]
    clades = [[diverge(a, clade_div) for a in ancestors] for _ in range(n_clades)]
    genomes = []
    for i in range(n_genomes):
        c = clades[i % n_clades]
        genomes.append([letters[diverge(seg, strain_div)].tobytes().decode() for seg in c])
    return genomes


def synthetic_taxa(n_taxa, genomes_per_taxon, seed=4, n_clades=10, clade_div=0.08, strain_div=0.02,
                   length_range=(10000, 30001)):
    """SURVEY.md section 8(d) config 4 (V-All shape): independent taxa, each with its own ancestor of
    a random length in [10 kb, 30 kb] and two-level divergence (clades, then strains) as in config 3.
    Returns a list of groups, each a list of genome strings (one sequence per genome)."""
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b'ACGT', dtype=np.uint8)

    def diverge(anc, div):
        mask = rng.random(len(anc)) < div
        shift = rng.integers(1, 4, len(anc))
        return np.where(mask, (anc + shift) % 4, anc)

    groups = []
    for _ in range(n_taxa):
        anc = rng.integers(0, 4, int(rng.integers(*length_range)))
        clades = [diverge(anc, clade_div) for _ in range(n_clades)]
        groups.append([letters[diverge(clades[i % n_clades], strain_div)].tobytes().decode()
                       for i in range(genomes_per_taxon)])
    return groups


def write_cli_inputs(tmp, spec):
    """FASTA files for a command-line case: one file per entry of `spec`; an entry is one generator call
    (n_genomes, length, divergence, seed) or a list of them (several families in one file)."""
    import os
    paths = []
    for gi, entry in enumerate(spec):
        calls = entry if isinstance(entry[0], (list, tuple)) else [entry]
        fn = os.path.join(str(tmp), 'g%d.fasta' % gi)
        with open(fn, 'w') as f:
            i = 0
            for n, length, div, seed in calls:
                for s in synthetic_genomes(n, length, div, seed):
                    f.write('>g%d\n%s\n' % (i, s))
                    i += 1
        paths.append(fn)
    return paths

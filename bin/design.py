#!/usr/bin/env python3
"""Design probes: same command line as the reference's bin/design.py (:448-980), with the
near-duplicate and set-cover filters running on the GPU (catch_b200).

The near-duplicate, set-cover and adapter filters and the coverage analysis run on the GPU
(catch_b200).  Flags that only configure parts of CATCH outside the accelerated path (NCBI download,
poly-A / N-expansion filters, --filter-from-fasta) are parsed for compatibility
and rejected with a clear message when used.
"""
import argparse
import logging
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from catch_b200 import __version__  # noqa: E402
from catch_b200 import coverage_analysis  # noqa: E402
from catch_b200.filter import (adapter_filter, duplicate_filter, near_duplicate_filter, probe_designer,  # noqa: E402
                                reverse_complement_filter, set_cover_filter)
from catch_b200.utils import seq_io  # noqa: E402

logger = logging.getLogger(__name__)

UNSUPPORTED = {
    'filter_from_fasta': '--filter-from-fasta', 'filter_polya': '--filter-polya',
    'expand_n': '--expand-n',
    'custom_hybridization_fn': '--custom-hybridization-fn',
    'custom_hybridization_fn_tolerant': '--custom-hybridization-fn-tolerant',
}


def main(args):
    logging.basicConfig(level=args.log_level, format='%(asctime)s - %(name)s - %(levelname)s - %(message)s')
    for attr, flag in UNSUPPORTED.items():
        if getattr(args, attr, None):
            raise SystemExit("%s configures a part of CATCH outside the GPU hot path and is not available "
                             "in this build" % flag)
    if args.cluster_and_design_separately and args.identify:              # bin/design.py:237-243
        raise Exception("Cannot use --cluster-and-design-separately with --identify, because clustering collapses "
                        "genome groupings into one")
    if args.cluster_from_fragments and not args.cluster_and_design_separately:
        raise Exception("Cannot use --cluster-from-fragments without also setting --cluster-and-design-separately")

    genomes_grouped = []
    genomes_grouped_names = []
    for ds in args.dataset:
        if ds.startswith('download:') or ds.startswith('collection:'):
            raise ValueError("Only FASTA files are accepted as input (no network access for 'download:')")
        if not os.path.isfile(ds):
            raise ValueError("Please check that the path to '%s' is valid" % ds)
        genomes_grouped.append(seq_io.read_genomes_from_fasta(ds))
        genomes_grouped_names.append(os.path.basename(ds))

    if args.limit_target_genomes and args.limit_target_genomes_randomly_with_replacement:
        raise Exception("Cannot --limit-target-genomes and --limit-target-genomes-randomly-with-replacement "
                        "at the same time")
    if args.limit_target_genomes:
        genomes_grouped = [g[:args.limit_target_genomes] for g in genomes_grouped]
    elif args.limit_target_genomes_randomly_with_replacement:
        k = args.limit_target_genomes_randomly_with_replacement
        genomes_grouped = [random.choices(g, k=k) for g in genomes_grouped]

    avoided = []
    for ag in args.avoid_genomes or []:
        if not os.path.isfile(ag):
            raise ValueError("Please check that the path to '%s' is valid" % ag)
        avoided.append(ag)

    if not args.lcf_thres:
        args.lcf_thres = args.probe_length
    if args.probe_stride > args.probe_length:
        logger.warning("PROBE_STRIDE (%d) is greater than PROBE_LENGTH (%d)", args.probe_stride, args.probe_length)
    if args.lcf_thres > args.probe_length:
        logger.warning("LCF_THRES (%d) is greater than PROBE_LENGTH (%d)", args.lcf_thres, args.probe_length)
    if args.kmer_probe_map_k:
        if args.kmer_probe_map_k > args.probe_length:
            raise Exception("KMER_PROBE_MAP_K (%d) exceeds PROBE_LENGTH (%d), which is not permitted" %
                            (args.kmer_probe_map_k, args.probe_length))
        k_scf = k_af = k_analyzer = args.kmer_probe_map_k
    else:
        # bin/design.py:199-205: 20 for the set cover and adapter filters, 10 for the (more sensitive) analyzer
        k_scf, k_af, k_analyzer = 20, 20, 10
    if args.add_adapters:
        if not (args.adapter_a or args.adapter_b):
            logger.warning("Adapter sequences will be added, but default sequences will be used; to provide "
                           "adapter sequences, use --adapter-a and --adapter-b")
    elif args.adapter_a or args.adapter_b:
        raise Exception("Adapter sequences were provided with --adapter-a and --adapter-b, but --add-adapters is "
                        "required to add adapter sequences onto the ends of probes")
    if args.small_seq_skip is not None and args.small_seq_min is not None:
        raise Exception("Both --small-seq-skip and --small-seq-min were specified, but both cannot be used together")

    # filter chain of bin/design.py:345-385 restricted to the hot path:
    # (NearDuplicate-Hamming | NearDuplicate-MinHash | Duplicate) -> SetCover
    filters = []
    if args.filter_with_lsh_hamming is not None and args.filter_with_lsh_minhash is not None:
        raise Exception("Cannot use both --filter-with-lsh-hamming and --filter-with-lsh-minhash")
    if args.filter_with_lsh_hamming is not None:
        filters.append(near_duplicate_filter.NearDuplicateFilterWithHammingDistance(
            args.filter_with_lsh_hamming, args.probe_length))
    elif args.filter_with_lsh_minhash is not None:
        filters.append(near_duplicate_filter.NearDuplicateFilterWithMinHash(args.filter_with_lsh_minhash))
    else:
        filters.append(duplicate_filter.DuplicateFilter())
    if not args.skip_set_cover:
        filters.append(set_cover_filter.SetCoverFilter(
            mismatches=args.mismatches, lcf_thres=args.lcf_thres,
            island_of_exact_match=args.island_of_exact_match,
            mismatches_tolerant=args.mismatches_tolerant, lcf_thres_tolerant=args.lcf_thres_tolerant,
            island_of_exact_match_tolerant=args.island_of_exact_match_tolerant,
            identify=args.identify, avoided_genomes=avoided, coverage=args.coverage,
            cover_extension=args.cover_extension, kmer_probe_map_k=k_scf))

    if args.add_adapters:                           # bin/design.py:343-365
        adapter_a = tuple(args.adapter_a) if args.adapter_a else ('ATACGCCATGCTGGGTCTCC', 'CGTACTTGGGAGTCGGCCAT')
        adapter_b = tuple(args.adapter_b) if args.adapter_b else ('AGGCCCTGGCTGCTGATATG', 'GACCTTTTGGGACAGCGGTG')
        filters.append(adapter_filter.AdapterFilter(adapter_a, adapter_b, mismatches=args.mismatches,
                                                    lcf_thres=args.lcf_thres,
                                                    island_of_exact_match=args.island_of_exact_match,
                                                    kmer_probe_map_k=k_af))

    if args.add_reverse_complements:               # bin/design.py:375-380
        filters.append(reverse_complement_filter.ReverseComplementFilter())

    if args.skip_set_cover:                         # bin/design.py:382-385 (the filter before the set cover)
        filter_before_scf = filters[0]
    if args.cluster_and_design_separately:          # bin/design.py:387-400
        cluster_kw = dict(cluster_threshold=args.cluster_and_design_separately,
                          cluster_merge_after=filter_before_scf if args.skip_set_cover else filters[1],
                          cluster_method=args.cluster_and_design_separately_method,
                          cluster_fragment_length=args.cluster_from_fragments)
    else:
        cluster_kw = {}
    pd = probe_designer.ProbeDesigner(genomes_grouped, filters, probe_length=args.probe_length,
                                      probe_stride=args.probe_stride, allow_small_seqs=args.small_seq_min,
                                      seq_length_to_skip=args.small_seq_skip, **cluster_kw)
    pd.design()
    seq_io.write_probe_fasta(pd.final_probes, args.output_probes)
    if (args.print_analysis or args.write_analysis_to_tsv or args.write_sliding_window_coverage or
            args.write_probe_map_counts_to_tsv):                          # bin/design.py:417-443
        analyzer = coverage_analysis.Analyzer(pd.final_probes, args.mismatches, args.lcf_thres, genomes_grouped,
                                              genomes_grouped_names,
                                              island_of_exact_match=args.island_of_exact_match,
                                              cover_extension=args.cover_extension, kmer_probe_map_k=k_analyzer,
                                              rc_too=bool(args.add_reverse_complements))
        analyzer.run()
        if args.write_analysis_to_tsv:
            analyzer.write_data_matrix_as_tsv(args.write_analysis_to_tsv)
        if args.write_sliding_window_coverage:
            analyzer.write_sliding_window_coverage(args.write_sliding_window_coverage)
        if args.write_probe_map_counts_to_tsv:
            analyzer.write_probe_map_counts(args.write_probe_map_counts_to_tsv)
        if args.print_analysis:
            analyzer.print_analysis()
    else:
        print(len(pd.final_probes))


def init_and_parse_args(args_type='basic', argv=None):
    ap = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument('dataset', nargs='+', help="FASTA file(s); each file is one grouping of target genomes")
    ap.add_argument('-o', '--output-probes', required=True, help="FASTA file to write the final probes to")
    ap.add_argument('--write-taxid-acc')
    ap.add_argument('-pl', '--probe-length', type=int, default=100)
    ap.add_argument('-ps', '--probe-stride', type=int, default=50)
    ap.add_argument('-m', '--mismatches', type=int, default={'basic': 0, 'large': 5}[args_type])
    ap.add_argument('-l', '--lcf-thres', type=int)
    ap.add_argument('--island-of-exact-match', type=int, default=0)
    ap.add_argument('--custom-hybridization-fn', nargs=2)

    def coverage(val):
        f = float(val)
        if 0 <= f <= 1:
            return f
        if f > 1 and f == int(f):
            return int(f)
        raise argparse.ArgumentTypeError("%s is an invalid coverage value" % val)
    ap.add_argument('-c', '--coverage', type=coverage, default=1.0)
    ap.add_argument('-e', '--cover-extension', type=int, default={'basic': 0, 'large': 50}[args_type])
    ap.add_argument('-i', '--identify', dest='identify', action='store_true')
    ap.add_argument('--avoid-genomes', nargs='+')
    ap.add_argument('-mt', '--mismatches-tolerant', type=int)
    ap.add_argument('-lt', '--lcf-thres-tolerant', type=int)
    ap.add_argument('--island-of-exact-match-tolerant', type=int, default=0)
    ap.add_argument('--custom-hybridization-fn-tolerant', nargs=2)
    ap.add_argument('--print-analysis', dest='print_analysis', action='store_true')
    ap.add_argument('--write-analysis-to-tsv')
    ap.add_argument('--write-sliding-window-coverage')
    ap.add_argument('--write-probe-map-counts-to-tsv')
    ap.add_argument('--filter-from-fasta')
    ap.add_argument('--skip-set-cover', dest='skip_set_cover', action='store_true')
    ap.add_argument('--add-adapters', dest='add_adapters', action='store_true')
    ap.add_argument('--adapter-a', nargs=2)
    ap.add_argument('--adapter-b', nargs=2)
    ap.add_argument('--filter-polya', nargs=2, type=int)
    ap.add_argument('--add-reverse-complements', dest='add_reverse_complements', action='store_true')
    ap.add_argument('--expand-n', nargs='?', type=int, default=None, const=3)
    ap.add_argument('--limit-target-genomes', type=int)
    ap.add_argument('--limit-target-genomes-randomly-with-replacement', type=int)

    def dissimilarity(val):
        f = float(val)
        if 0 < f <= 0.5:
            return f
        raise argparse.ArgumentTypeError("%s is an invalid average nucleotide dissimilarity" % val)
    ap.add_argument('--cluster-and-design-separately', type=dissimilarity,
                    default={'basic': None, 'large': 0.15}[args_type])
    ap.add_argument('--cluster-and-design-separately-method', choices=['choose', 'simple', 'hierarchical'],
                    default='choose')
    ap.add_argument('--cluster-from-fragments', type=int, default={'basic': None, 'large': 50000}[args_type])
    ap.add_argument('--filter-with-lsh-hamming', type=int)

    def jaccard(val):
        f = float(val)
        if 0.0 <= f <= 1.0:
            return f
        raise argparse.ArgumentTypeError("%s is an invalid Jaccard distance" % val)
    ap.add_argument('--filter-with-lsh-minhash', type=jaccard, default={'basic': None, 'large': 0.6}[args_type])
    ap.add_argument('--small-seq-skip', type=int)
    ap.add_argument('--small-seq-min', type=int)

    def n_proc(val):
        i = int(val)
        if i >= 1:
            return i
        raise argparse.ArgumentTypeError("MAX_NUM_PROCESSES must be an int >= 1")
    ap.add_argument('--max-num-processes', type=n_proc, default=None,
                    help="accepted for compatibility; the GPU path has no process pools")
    ap.add_argument('--kmer-probe-map-k', type=int)
    ap.add_argument('--use-native-dict-when-finding-tolerant-coverage',
                    dest='use_native_dict_when_finding_tolerant_coverage', action='store_true')
    ap.add_argument('--ncbi-api-key')
    ap.add_argument('--debug', dest='log_level', action='store_const', const=logging.DEBUG, default=logging.WARNING)
    ap.add_argument('--verbose', dest='log_level', action='store_const', const=logging.INFO)
    ap.add_argument('-V', '--version', action='version', version='catch_b200 ' + __version__)
    args = ap.parse_args(argv)
    args.args_type = args_type
    return args


if __name__ == '__main__':
    main(init_and_parse_args('basic'))

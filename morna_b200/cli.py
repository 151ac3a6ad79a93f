"""`morna.py index` / `morna.py search` command line of the reference
(morna.py:876-1054, dispatch :1338-1484) in front of the B200 kernels.

Kept: every flag of the two subcommands, stdin for the query, the
"rank.<TAB>id[<TAB>distance][<TAB>metadata]" result lines with Python 2's float
formatting, progress messages, the two extra stdout lines of -q.
Different on purpose: without -e the reference asks Annoy for approximate
neighbours; there is no forest here -- the approximate mode is the tensor-core
scoring pass without the exact re-rank (MornaSearch.search_nn; --search-k,
--n-trees, -b are accepted and ignored), -e is the exact search.  `-q ID -e` runs the exact
search with the stored row as the query (the reference ignores -e after -q).
-m works as in the reference (index: metadata file -> <basename>.meta.mor; search: keywords joined to the results).
The `junctions` subcommand and the convergence back-off loop are outside the hot path and exit with a message.
"""
import argparse
import sys

_help_intro = "morna (B200-native hot path): index and exactly search junction feature vectors"


def py2_str(v):
    """Python 2 str(): floats print 12 significant digits (morna.py:126 uses str())."""
    if isinstance(v, float):
        s = "%.12g" % v
        if s.lstrip("-").isdigit():
            s += ".0"
        return s
    if isinstance(v, tuple) and len(v) == 1 and isinstance(v[0], str):
        return "(u" + repr(v[0]) + ",)"          # a metadata row as Python 2 prints the sqlite tuple of one unicode string
    return str(v)


def results_output(results, out=None):
    """morna.py:116-127."""
    out = out or sys.stdout
    for i in range(len(results[0])):
        out.write(str(i + 1) + ".")
        for column in results:
            out.write("\t" + py2_str(column[i]))
        out.write("\n")


def add_search_parameters(sub):
    sub.add_argument("-x", "--basename", metavar="<idx>", type=str, required=True,
                     help="path to junction index basename for search")
    sub.add_argument("-v", "--verbose", action="store_const", const=True, default=False, help="be talkative")
    sub.add_argument("--search-k", metavar="<int>", type=int, required=False, default=100,
                     help="accepted for compatibility; the exact search ignores it")
    sub.add_argument("-f", "--format", metavar="<choice>", type=str, required=False, default="sam",
                     help="one of {sam, bed, raw}")
    sub.add_argument("-d", "--distances", action="store_const", const=True, default=False,
                     help="include distances to nearest neighbors")
    sub.add_argument("-m", "--metadata", action="store_const", const=True, default=False,
                     help="display results mapped to metadata")
    sub.add_argument("-c", "--convergence-backoff", metavar="<int>", type=int, required=False, default=None,
                     help="not supported (the reference calls it non-functional)")
    sub.add_argument("-ch", "--checkpoint", metavar="<int>", type=int, required=False, default=0,
                     help="not supported")
    sub.add_argument("-q", "--query-id", metavar="<int>", type=int, required=False, default=None,
                     help="search for nearest neighbors of the indexed sample with this sample id")
    sub.add_argument("-e", "--exact", action="store_const", const=True, default=False,
                     help="exact nearest neighbor search (always on in this implementation)")
    sub.add_argument("-r", "--results", metavar="<int>", type=int, required=False, default=20,
                     help="the number of nearest neighbor results to return")
    sub.add_argument("-rl", "--rawlist", action="store_const", const=True, default=False,
                     help="regurgitate junction list for input sample instead of performing search")


def build_parser():
    parser = argparse.ArgumentParser(description=_help_intro)
    subs = parser.add_subparsers(dest="subparser_name",
                                 help='subcommands; add "-h" or "--help" after a subcommand for its parameters')
    index_parser = subs.add_parser("index", help="creates a morna index")
    search_parser = subs.add_parser("search", help="searches a morna index")
    index_parser.add_argument("--intropolis", metavar="<file>", type=str, required=True,
                              help="path to (gzipped) file recording junctions across samples in intropolis format")
    index_parser.add_argument("-x", "--basename", metavar="<str>", type=str, required=False, default="morna",
                              help="basename path of junction index files to create")
    index_parser.add_argument("--features", metavar="<int>", type=int, required=False, default=3000,
                              help="dimension of feature space")
    index_parser.add_argument("--n-trees", metavar="<int>", type=int, required=False, default=200,
                              help="accepted for compatibility; no Annoy forest is built")
    index_parser.add_argument("-s", "--sample-count", metavar="<int>", type=int, required=False, default=None,
                              help="optionally specify number of unique samples to speed indexing")
    index_parser.add_argument("-t", "--sample-threshold", metavar="<int>", type=int, required=False, default=100,
                              help="minimum number of samples in which a junction should appear")
    index_parser.add_argument("-b", "--buffer-size", metavar="<int>", type=int, required=False, default=1024,
                              help="accepted for compatibility (junction database buffer)")
    index_parser.add_argument("-v", "--verbose", action="store_const", const=True, default=False,
                              help="be talkative")
    index_parser.add_argument("-m", "--metafile", metavar="<file>", type=str, required=False, default=None,
                              help="metadata file: sample id then keywords per line; stored in <basename>.meta.mor for search -m")
    index_parser.add_argument("--no-junction-shards", action="store_const", const=True, default=False,
                              help="do not write the <basename>.shXX.junc.mor shards (only `junctions` reads them)")
    add_search_parameters(search_parser)
    junctions_parser = subs.add_parser("junctions", help="extracts the junctions of a query's nearest neighbors")
    add_search_parameters(junctions_parser)                      # morna.py:1023-1054
    junctions_parser.add_argument("-i", "--index", metavar="<idx>", type=str, required=True,
                                  help="index basename or directory of the aligner (kept for compatibility)")
    junctions_parser.add_argument("-p1", "--pass1-sam", metavar="<sam>", type=str, required=False, default="pass1.sam",
                                  help="filename for first pass alignment file output by aligner")
    junctions_parser.add_argument("--junction-filter", type=str, required=False, default=".05,5",
                                  help="retain junctions found in at least {first part} of the result samples, or "
                                       "with at least {second part} coverage in any one result sample")
    junctions_parser.add_argument("--junction-file", type=str, metavar="<gz>", required=True,
                                  help="gzipped file with junction rows in the order of the file the index was made from")
    junctions_parser.add_argument("-sf", "--splicefile", type=str, metavar="<file>", required=True,
                                  help="output intropolis-like file with the retained junctions")
    return parser


def main(argv=None, stdin=None, stdout=None, stderr=None):
    stdin, stdout, stderr = stdin or sys.stdin, stdout or sys.stdout, stderr or sys.stderr
    parser = build_parser()
    args = parser.parse_args(argv)
    if args.subparser_name is None:
        parser.print_help(stderr)
        return 2
    if args.subparser_name == "index":
        from .index import go_index
        go_index(args.intropolis, args.basename, args.features, args.n_trees, args.sample_count,
                 args.sample_threshold, args.buffer_size, args.verbose, args.metafile, out=stdout,
                 junction_shards=not args.no_junction_shards)
        return 0

    from . import parse
    from .search import MornaSearch
    if args.convergence_backoff:
        stderr.write("convergence back-off is not supported (see README of the reference: non-functional)\n")
        return 2
    opened = None
    if args.subparser_name == "junctions":                      # morna.py:1351-1353
        args.format = "sam"
        stdin = opened = open(args.pass1_sam)
    searcher = MornaSearch(basename=args.basename)
    if args.query_id is not None:                               # morna.py:1358-1365
        results = searcher.search_member_n(args.query_id, args.results, args.search_k,
                                           include_distances=args.distances, meta_db=args.metadata, out=stdout)
        results_output(results, stdout)
        return 0
    if args.format == "sam":                                    # :1367-1373
        junctions = parse.junctions_from_sam_stream(stdin)
    elif args.format == "bed":
        junctions = parse.junctions_from_bed_stream(stdin)
    else:
        assert args.format == "raw"
        junctions = parse.junctions_from_raw_stream(stdin)
    if args.rawlist:                                            # :1374-1377
        for junction in junctions:
            stdout.write(str(junction) + "\n")
        return 0
    for i, junction in enumerate(junctions):                    # :1456-1473
        if args.verbose and i % 1000 == 0:
            stderr.write(str(i) + " junctions into query sample\r")
            stderr.flush()
        if " ".join(str(t) for t in junction[:3]) in searcher.sample_frequencies:
            searcher.update_query(junction)
    searcher.finalize_query()
    if args.verbose:
        stderr.write("\n")
    if args.exact:                                              # :1477-1483
        results = searcher.exact_search_nn(args.results, include_distances=args.distances, meta_db=args.metadata)
    else:
        results = searcher.search_nn(args.results, args.search_k, include_distances=args.distances, meta_db=args.metadata)
    results_output(results, stdout)
    if args.subparser_name == "junctions":                      # :1486-1632
        from .junctions import go_junctions
        go_junctions(args, searcher, results, stdout, stderr)
        opened.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

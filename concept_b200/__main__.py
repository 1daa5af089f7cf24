"""`python -m concept_b200 -p <parameter file> [-c "name = value" ...]` — the particle-only counterpart of the
reference's `concept -p param` launcher (the bash launcher/installer themselves are out of scope):
load the parameter file, realise or load the initial conditions, run the time loop, write the requested outputs.
Under torchrun one process drives one GPU (x-slab decomposition)."""
import argparse
import sys

from . import commons, communication, main, mesh


def cli(argv=None):
    parser = argparse.ArgumentParser(prog='python -m concept_b200', description=__doc__)
    parser.add_argument('-p', '--params', default='', help='CO*N*CEPT parameter file')
    parser.add_argument('-u', '--utility', nargs='+', metavar=('NAME', 'PATH'), default=None,
                        help='run a utility instead of a simulation: "powerspec SNAPSHOT…" or "info SNAPSHOT…" (util/ of the reference)')
    parser.add_argument('-c', '--command-line-params', action='append', default=[],
                        help='extra parameter assignments, executed after the file (repeatable)')
    parser.add_argument('--max-steps', type=int, default=None, help='stop after this many base time steps')
    args = parser.parse_args(argv)
    communication.init()
    commons.verbose = True
    extra = '\n'.join(args.command_line_params)
    if args.utility:
        name, paths = args.utility[0], args.utility[1:]
        if name not in ('powerspec', 'info') or not paths:
            parser.error('utilities: powerspec SNAPSHOT…, info SNAPSHOT…')
        try:
            for path in paths:
                if name == 'info':
                    main.utility_info(path)
                else:
                    commons.masterprint('power spectrum written to', main.utility_powerspec(path, args.params, extra))
        finally:
            mesh.free_contexts()
        communication.finalize()
        return 0
    if not args.params:
        parser.error('a parameter file (-p) is needed for a simulation')
    try:
        components = main.run(args.params, '\n'.join(args.command_line_params), max_steps=args.max_steps)
    except BaseException:
        mesh.free_contexts()
        raise
    mesh.free_contexts()
    commons.masterprint(f'concept_b200 run finished: {len(components)} component(s), a = {commons.universals.a:.6g}')
    communication.finalize()
    return 0


if __name__ == '__main__':
    sys.exit(cli())

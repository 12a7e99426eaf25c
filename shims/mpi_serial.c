/*
 * mpi_serial.c -- ENVIRONMENT SHIM (see shims/README.md).
 *
 * One-rank implementation of the MPI subset the reference (SPARC) calls.
 * Every collective degenerates to a local copy; point-to-point traffic can
 * only be rank 0 -> rank 0 and is matched through a small in-process queue
 * (needed by D2D, /root/reference/src/parallelization.c:2253, sends :2367,
 * receives :2437).  Communicators carry just enough state (Cartesian dims and
 * periodicity) to answer MPI_Cart_* queries.
 */
#include "mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ comms */
#define MAX_COMMS 4096
typedef struct {
    int used;
    int ndims;
    int dims[8];
    int periods[8];
} comm_t;

static comm_t g_comm[MAX_COMMS] = {
    {0, 0, {0}, {0}}, /* MPI_COMM_NULL */
    {1, 0, {0}, {0}}, /* MPI_COMM_WORLD */
    {1, 0, {0}, {0}}, /* MPI_COMM_SELF */
};

static void die(const char *msg)
{
    fprintf(stderr, "[mpi_serial] fatal: %s\n", msg);
    abort();
}

static MPI_Comm comm_new(void)
{
    for (int i = 3; i < MAX_COMMS; i++) {
        if (!g_comm[i].used) {
            memset(&g_comm[i], 0, sizeof(comm_t));
            g_comm[i].used = 1;
            return i;
        }
    }
    die("communicator table exhausted");
    return MPI_COMM_NULL;
}

static void comm_check(MPI_Comm c, const char *who)
{
    if (c <= 0 || c >= MAX_COMMS || !g_comm[c].used) {
        fprintf(stderr, "[mpi_serial] %s called on invalid/null communicator %d\n", who, c);
        abort();
    }
}

/* -------------------------------------------------------------- datatypes */
#define MAX_TYPES 256
static size_t g_user_type_size[MAX_TYPES];
static int g_user_type_used[MAX_TYPES];

static size_t type_size(MPI_Datatype t)
{
    switch (t) {
    case MPI_CHAR: return 1;
    case MPI_INT: return sizeof(int);
    case MPI_DOUBLE: return sizeof(double);
    case MPI_DOUBLE_COMPLEX: return 2 * sizeof(double);
    case MPI_PACKED: return 1;
    case MPI_LONG: return sizeof(long);
    case MPI_FLOAT: return sizeof(float);
    case MPI_UNSIGNED: return sizeof(unsigned);
    case MPI_BYTE: return 1;
    default:
        if (t >= 16 && t < 16 + MAX_TYPES && g_user_type_used[t - 16]) return g_user_type_size[t - 16];
    }
    die("unknown datatype");
    return 0;
}

static void copy_elems(const void *src, void *dst, long count, MPI_Datatype t)
{
    if (src == MPI_IN_PLACE || src == dst || count <= 0) return;
    memmove(dst, src, (size_t)count * type_size(t));
}

/* ---------------------------------------------------------------- basics */
int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int errorcode)
{
    (void)comm;
    fprintf(stderr, "[mpi_serial] MPI_Abort(%d)\n", errorcode);
    exit(errorcode ? errorcode : 1);
}
double MPI_Wtime(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Comm_rank(MPI_Comm comm, int *rank) { comm_check(comm, "MPI_Comm_rank"); *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { comm_check(comm, "MPI_Comm_size"); *size = 1; return MPI_SUCCESS; }

int MPI_Comm_split(MPI_Comm comm, int color, int key, MPI_Comm *newcomm)
{
    (void)key;
    comm_check(comm, "MPI_Comm_split");
    *newcomm = (color == MPI_UNDEFINED) ? MPI_COMM_NULL : comm_new();
    return MPI_SUCCESS;
}

int MPI_Comm_free(MPI_Comm *comm)
{
    if (*comm >= 3 && *comm < MAX_COMMS) g_comm[*comm].used = 0;
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

MPI_Fint MPI_Comm_c2f(MPI_Comm comm) { return comm; }

/* groups: handle 1 = empty, 2 = {rank 0} */
#define GROUP_SELF 2
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) { comm_check(comm, "MPI_Comm_group"); *group = GROUP_SELF; return MPI_SUCCESS; }
int MPI_Group_incl(MPI_Group group, int n, const int ranks[], MPI_Group *newgroup)
{
    *newgroup = MPI_GROUP_EMPTY;
    if (group == GROUP_SELF)
        for (int i = 0; i < n; i++) if (ranks[i] == 0) *newgroup = GROUP_SELF;
    return MPI_SUCCESS;
}
int MPI_Group_excl(MPI_Group group, int n, const int ranks[], MPI_Group *newgroup)
{
    *newgroup = group;
    for (int i = 0; i < n; i++) if (ranks[i] == 0) *newgroup = MPI_GROUP_EMPTY;
    return MPI_SUCCESS;
}
int MPI_Group_free(MPI_Group *group) { *group = MPI_GROUP_NULL; return MPI_SUCCESS; }
int MPI_Group_translate_ranks(MPI_Group g1, int n, const int ranks1[], MPI_Group g2, int ranks2[])
{
    (void)g1;
    for (int i = 0; i < n; i++) ranks2[i] = (g2 == GROUP_SELF && ranks1[i] == 0) ? 0 : MPI_UNDEFINED;
    return MPI_SUCCESS;
}
int MPI_Comm_create_group(MPI_Comm comm, MPI_Group group, int tag, MPI_Comm *newcomm)
{
    (void)tag;
    comm_check(comm, "MPI_Comm_create_group");
    *newcomm = (group == GROUP_SELF) ? comm_new() : MPI_COMM_NULL;
    return MPI_SUCCESS;
}
int MPI_Intercomm_create(MPI_Comm local_comm, int local_leader, MPI_Comm peer_comm,
                         int remote_leader, int tag, MPI_Comm *newintercomm)
{
    (void)local_comm; (void)local_leader; (void)peer_comm; (void)remote_leader; (void)tag;
    /* an inter-communicator needs a non-empty remote group: impossible with one rank */
    *newintercomm = MPI_COMM_NULL;
    die("MPI_Intercomm_create reached with a single rank");
    return MPI_ERR_OTHER;
}

/* ------------------------------------------------------------ topologies */
int MPI_Cart_create(MPI_Comm comm_old, int ndims, const int dims[], const int periods[],
                    int reorder, MPI_Comm *comm_cart)
{
    (void)reorder;
    comm_check(comm_old, "MPI_Cart_create");
    long np = 1;
    for (int i = 0; i < ndims; i++) np *= dims[i];
    if (np > 1) die("MPI_Cart_create with more than one process");
    if (np < 1 || ndims > 8) { *comm_cart = MPI_COMM_NULL; return MPI_SUCCESS; }
    MPI_Comm c = comm_new();
    g_comm[c].ndims = ndims;
    for (int i = 0; i < ndims; i++) { g_comm[c].dims[i] = dims[i]; g_comm[c].periods[i] = periods[i]; }
    *comm_cart = c;
    return MPI_SUCCESS;
}
int MPI_Cart_get(MPI_Comm comm, int maxdims, int dims[], int periods[], int coords[])
{
    comm_check(comm, "MPI_Cart_get");
    for (int i = 0; i < maxdims; i++) {
        dims[i] = (i < g_comm[comm].ndims) ? g_comm[comm].dims[i] : 1;
        periods[i] = (i < g_comm[comm].ndims) ? g_comm[comm].periods[i] : 0;
        coords[i] = 0;
    }
    return MPI_SUCCESS;
}
int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int coords[])
{
    (void)rank;
    comm_check(comm, "MPI_Cart_coords");
    for (int i = 0; i < maxdims; i++) coords[i] = 0;
    return MPI_SUCCESS;
}
int MPI_Cart_rank(MPI_Comm comm, const int coords[], int *rank)
{
    comm_check(comm, "MPI_Cart_rank");
    /* non-periodic out-of-range coordinates are erroneous in MPI; callers guard them */
    (void)coords;
    *rank = 0;
    return MPI_SUCCESS;
}
int MPI_Cart_sub(MPI_Comm comm, const int remain_dims[], MPI_Comm *newcomm)
{
    comm_check(comm, "MPI_Cart_sub");
    MPI_Comm c = comm_new();
    int nd = 0;
    for (int i = 0; i < g_comm[comm].ndims; i++) {
        if (remain_dims[i]) {
            g_comm[c].dims[nd] = g_comm[comm].dims[i];
            g_comm[c].periods[nd] = g_comm[comm].periods[i];
            nd++;
        }
    }
    g_comm[c].ndims = nd;
    *newcomm = c;
    return MPI_SUCCESS;
}
int MPI_Dist_graph_create_adjacent(MPI_Comm comm_old, int indegree, const int sources[],
                                   const int sourceweights[], int outdegree,
                                   const int destinations[], const int destweights[],
                                   MPI_Info info, int reorder, MPI_Comm *comm_dist_graph)
{
    (void)indegree; (void)sources; (void)sourceweights; (void)outdegree; (void)destinations;
    (void)destweights; (void)info; (void)reorder;
    comm_check(comm_old, "MPI_Dist_graph_create_adjacent");
    *comm_dist_graph = comm_new();
    return MPI_SUCCESS;
}

/* ------------------------------------------------------------ collectives */
int MPI_Barrier(MPI_Comm comm) { comm_check(comm, "MPI_Barrier"); return MPI_SUCCESS; }
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm)
{
    (void)buffer; (void)count; (void)datatype; (void)root;
    comm_check(comm, "MPI_Bcast");
    return MPI_SUCCESS;
}
int MPI_Ibcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm,
               MPI_Request *request)
{
    *request = MPI_REQUEST_NULL;
    return MPI_Bcast(buffer, count, datatype, root, comm);
}
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype,
                  MPI_Op op, MPI_Comm comm)
{
    (void)op;
    comm_check(comm, "MPI_Allreduce");
    copy_elems(sendbuf, recvbuf, count, datatype);
    return MPI_SUCCESS;
}
int MPI_Iallreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype,
                   MPI_Op op, MPI_Comm comm, MPI_Request *request)
{
    *request = MPI_REQUEST_NULL;
    return MPI_Allreduce(sendbuf, recvbuf, count, datatype, op, comm);
}
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype, MPI_Op op,
               int root, MPI_Comm comm)
{
    (void)root;
    return MPI_Allreduce(sendbuf, recvbuf, count, datatype, op, comm);
}
int MPI_Ireduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype datatype,
                MPI_Op op, int root, MPI_Comm comm, MPI_Request *request)
{
    *request = MPI_REQUEST_NULL;
    return MPI_Reduce(sendbuf, recvbuf, count, datatype, op, root, comm);
}
int MPI_Allgather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                  int recvcount, MPI_Datatype recvtype, MPI_Comm comm)
{
    (void)recvcount; (void)recvtype;
    comm_check(comm, "MPI_Allgather");
    copy_elems(sendbuf, recvbuf, sendcount, sendtype);
    return MPI_SUCCESS;
}
int MPI_Allgatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                   const int recvcounts[], const int displs[], MPI_Datatype recvtype,
                   MPI_Comm comm)
{
    comm_check(comm, "MPI_Allgatherv");
    if (sendbuf != MPI_IN_PLACE) {
        (void)recvcounts;
        copy_elems(sendbuf, (char *)recvbuf + (size_t)displs[0] * type_size(recvtype), sendcount, sendtype);
    }
    return MPI_SUCCESS;
}
int MPI_Gather(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
               int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
    (void)root;
    return MPI_Allgather(sendbuf, sendcount, sendtype, recvbuf, recvcount, recvtype, comm);
}
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, void *recvbuf,
                const int recvcounts[], const int displs[], MPI_Datatype recvtype, int root,
                MPI_Comm comm)
{
    (void)root;
    return MPI_Allgatherv(sendbuf, sendcount, sendtype, recvbuf, recvcounts, displs, recvtype, comm);
}
int MPI_Scatterv(const void *sendbuf, const int sendcounts[], const int displs[],
                 MPI_Datatype sendtype, void *recvbuf, int recvcount, MPI_Datatype recvtype,
                 int root, MPI_Comm comm)
{
    (void)root; (void)recvcount; (void)recvtype;
    comm_check(comm, "MPI_Scatterv");
    if (recvbuf != MPI_IN_PLACE)
        copy_elems((const char *)sendbuf + (size_t)displs[0] * type_size(sendtype), recvbuf, sendcounts[0], sendtype);
    return MPI_SUCCESS;
}
int MPI_Alltoallv(const void *sendbuf, const int sendcounts[], const int sdispls[],
                  MPI_Datatype sendtype, void *recvbuf, const int recvcounts[],
                  const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm)
{
    (void)recvcounts;
    comm_check(comm, "MPI_Alltoallv");
    copy_elems((const char *)sendbuf + (size_t)sdispls[0] * type_size(sendtype),
               (char *)recvbuf + (size_t)rdispls[0] * type_size(recvtype), sendcounts[0], sendtype);
    return MPI_SUCCESS;
}
int MPI_Ineighbor_alltoallv(const void *sendbuf, const int sendcounts[], const int sdispls[],
                            MPI_Datatype sendtype, void *recvbuf, const int recvcounts[],
                            const int rdispls[], MPI_Datatype recvtype, MPI_Comm comm,
                            MPI_Request *request)
{
    (void)sendbuf; (void)sendcounts; (void)sdispls; (void)sendtype; (void)recvbuf;
    (void)recvcounts; (void)rdispls; (void)recvtype; (void)comm; (void)request;
    /* the reference only exchanges halos when the domain is split (nproc > 1),
       lapVecRoutines.c:387,1009 -- never with one rank */
    die("MPI_Ineighbor_alltoallv reached with a single rank");
    return MPI_ERR_OTHER;
}

/* ------------------------------------------------- self point-to-point */
typedef struct msg_s {
    MPI_Comm comm;
    int tag;
    size_t bytes;
    void *data;       /* owned copy (pending send) */
    void *recvbuf;    /* destination (pending recv) */
    struct msg_s *next;
} msg_t;
static msg_t *g_sends = NULL, *g_recvs = NULL;

static msg_t *queue_take(msg_t **head, MPI_Comm comm, int tag)
{
    for (msg_t **pp = head; *pp; pp = &(*pp)->next) {
        if ((*pp)->comm == comm && ((*pp)->tag == tag || tag == MPI_ANY_TAG)) {
            msg_t *m = *pp;
            *pp = m->next;
            return m;
        }
    }
    return NULL;
}
static void queue_push(msg_t **head, msg_t *m)
{
    m->next = NULL;
    while (*head) head = &(*head)->next;
    *head = m;
}

int MPI_Isend(const void *buf, int count, MPI_Datatype datatype, int dest, int tag,
              MPI_Comm comm, MPI_Request *request)
{
    if (request) *request = MPI_REQUEST_NULL;
    if (dest == MPI_PROC_NULL) return MPI_SUCCESS;
    comm_check(comm, "MPI_Isend");
    if (dest != 0) die("send to a rank other than 0");
    size_t bytes = (size_t)count * type_size(datatype);
    msg_t *r = queue_take(&g_recvs, comm, tag);
    if (r) {
        if (r->bytes < bytes) die("self-send larger than posted receive");
        memcpy(r->recvbuf, buf, bytes);
        free(r);
        return MPI_SUCCESS;
    }
    msg_t *m = (msg_t *)calloc(1, sizeof(msg_t));
    m->comm = comm; m->tag = tag; m->bytes = bytes;
    m->data = malloc(bytes ? bytes : 1);
    memcpy(m->data, buf, bytes);
    queue_push(&g_sends, m);
    return MPI_SUCCESS;
}
int MPI_Irecv(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm,
              MPI_Request *request)
{
    if (request) *request = MPI_REQUEST_NULL;
    if (source == MPI_PROC_NULL) return MPI_SUCCESS;
    comm_check(comm, "MPI_Irecv");
    if (source != 0 && source != MPI_ANY_SOURCE) die("receive from a rank other than 0");
    size_t bytes = (size_t)count * type_size(datatype);
    msg_t *s = queue_take(&g_sends, comm, tag);
    if (s) {
        if (s->bytes > bytes) die("self-send larger than receive buffer");
        memcpy(buf, s->data, s->bytes);
        free(s->data);
        free(s);
        return MPI_SUCCESS;
    }
    msg_t *m = (msg_t *)calloc(1, sizeof(msg_t));
    m->comm = comm; m->tag = tag; m->bytes = bytes; m->recvbuf = buf;
    queue_push(&g_recvs, m);
    return MPI_SUCCESS;
}
int MPI_Send(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm)
{
    return MPI_Isend(buf, count, datatype, dest, tag, comm, NULL);
}
int MPI_Recv(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm,
             MPI_Status *status)
{
    if (status) { status->MPI_SOURCE = 0; status->MPI_TAG = tag; status->MPI_ERROR = MPI_SUCCESS; }
    if (source == MPI_PROC_NULL) return MPI_SUCCESS;
    comm_check(comm, "MPI_Recv");
    msg_t *s = queue_take(&g_sends, comm, tag);
    if (!s) die("blocking MPI_Recv with no matching self-send would deadlock");
    memcpy(buf, s->data, s->bytes);
    (void)count; (void)datatype;
    free(s->data);
    free(s);
    return MPI_SUCCESS;
}
int MPI_Sendrecv(const void *sendbuf, int sendcount, MPI_Datatype sendtype, int dest,
                 int sendtag, void *recvbuf, int recvcount, MPI_Datatype recvtype, int source,
                 int recvtag, MPI_Comm comm, MPI_Status *status)
{
    MPI_Isend(sendbuf, sendcount, sendtype, dest, sendtag, comm, NULL);
    return MPI_Recv(recvbuf, recvcount, recvtype, source, recvtag, comm, status);
}
int MPI_Wait(MPI_Request *request, MPI_Status *status)
{
    (void)status;
    if (g_recvs) die("MPI_Wait with an unmatched self-receive pending");
    if (request) *request = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status stats[])
{
    (void)stats;
    if (g_recvs) die("MPI_Waitall with an unmatched self-receive pending");
    for (int i = 0; i < count; i++) reqs[i] = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
int MPI_Test(MPI_Request *request, int *flag, MPI_Status *status)
{
    (void)status;
    if (request) *request = MPI_REQUEST_NULL;
    *flag = 1;
    return MPI_SUCCESS;
}

/* ---------------------------------------------------- datatypes / packing */
int MPI_Get_address(const void *location, MPI_Aint *address)
{
    *address = (MPI_Aint)location;
    return MPI_SUCCESS;
}
int MPI_Type_create_struct(int count, const int blocklengths[], const MPI_Aint displs[],
                           const MPI_Datatype types[], MPI_Datatype *newtype)
{
    size_t extent = 0;
    for (int i = 0; i < count; i++) {
        size_t end = (size_t)displs[i] + (size_t)blocklengths[i] * type_size(types[i]);
        if (end > extent) extent = end;
    }
    for (int i = 0; i < MAX_TYPES; i++) {
        if (!g_user_type_used[i]) {
            g_user_type_used[i] = 1;
            g_user_type_size[i] = extent;
            *newtype = 16 + i;
            return MPI_SUCCESS;
        }
    }
    die("datatype table exhausted");
    return MPI_ERR_OTHER;
}
int MPI_Type_commit(MPI_Datatype *datatype) { (void)datatype; return MPI_SUCCESS; }
int MPI_Type_free(MPI_Datatype *datatype)
{
    if (*datatype >= 16 && *datatype < 16 + MAX_TYPES) g_user_type_used[*datatype - 16] = 0;
    *datatype = MPI_DATATYPE_NULL;
    return MPI_SUCCESS;
}
int MPI_Pack(const void *inbuf, int incount, MPI_Datatype datatype, void *outbuf, int outsize,
             int *position, MPI_Comm comm)
{
    (void)comm;
    size_t bytes = (size_t)incount * type_size(datatype);
    if ((size_t)*position + bytes > (size_t)outsize) die("MPI_Pack overflow");
    memcpy((char *)outbuf + *position, inbuf, bytes);
    *position += (int)bytes;
    return MPI_SUCCESS;
}
int MPI_Unpack(const void *inbuf, int insize, int *position, void *outbuf, int outcount,
               MPI_Datatype datatype, MPI_Comm comm)
{
    (void)comm;
    size_t bytes = (size_t)outcount * type_size(datatype);
    if ((size_t)*position + bytes > (size_t)insize) die("MPI_Unpack overflow");
    memcpy(outbuf, (const char *)inbuf + *position, bytes);
    *position += (int)bytes;
    return MPI_SUCCESS;
}

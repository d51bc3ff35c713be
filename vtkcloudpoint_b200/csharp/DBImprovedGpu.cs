// DBImprovedGpu.cs -- drop-in for vtkPointCloud.DBImproved (BaseClass/DBImproved.cs:8-116): same public
// fields, same dbscan(List<Point3D>, double, int) signature and the same mutation contract, computed by
// libvpc on the GPU.  Swap `new DBImproved()` for `new DBImprovedGpu()` at FrmMain.cs:1508, :2785 and
// Tools.cs:591.  Source only (no .NET toolchain in the build image).
using System;
using System.Collections.Generic;

namespace vtkPointCloud
{
    public class DBImprovedGpu : IDisposable
    {
        public int clusterAmount = 0;   // DBImproved.cs:10
        public int pointsAmount = 0;    // DBImproved.cs:11
        public int cf = 0;              // DBImproved.cs:13 -- may be pre-seeded by the caller (FrmMain.cs:1509)
        private IntPtr ctx;

        public DBImprovedGpu() { NativeMethods.Check(IntPtr.Zero, NativeMethods.vpc_create(out ctx, null, 0)); }

        public void dbscan(List<Point3D> lst, double e, int minPts)
        {
            int n = lst.Count;
            double[] mx = new double[n], my = new double[n];
            for (int i = 0; i < n; i++) { mx[i] = lst[i].motor_x; my[i] = lst[i].motor_y; }   // getDisP reads only these (DBImproved.cs:16-17)
            int[] cid = new int[n]; byte[] key = new byte[n], cls = new byte[n];
            int amount;
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_l1_2d(ctx, mx, my, n, e, minPts, cf, cid, key, cls, out amount));
            for (int i = 0; i < n; i++)
            {
                // The C# never clears flags itself; its callers reset clusterId/isClassed first (FrmMain.cs:1219-1223,
                // 1512-1515; Tools.cs:584-590), so writing all three fields reproduces the post-state exactly.
                lst[i].clusterId = cid[i];
                lst[i].isClassed = cls[i] != 0;
                if (key[i] != 0) lst[i].isKeyPoint = true;   // isKeyPoint is only ever set, never cleared (DBImproved.cs:49)
            }
            pointsAmount += n;        // DBImproved.cs:99
            cf = amount;              // DBImproved.cs:107
            clusterAmount = amount;   // DBImproved.cs:112
        }

        // All cells of the blocked clustering at once: replaces the ThreadPool loop of DoWork3 (FrmMain.cs:1356-1359).
        // Returns clusterAmount per cell; cluster ids written to the points are cell-local like StartCode's.
        public int[] dbscanCells(List<Point3D>[] cells, double e, int minPts)
        {
            long[] off = new long[cells.Length + 1];
            for (int k = 0; k < cells.Length; k++) off[k + 1] = off[k] + (cells[k] == null ? 0 : cells[k].Count);
            int n = (int)off[cells.Length];
            double[] mx = new double[n], my = new double[n];
            for (int k = 0, p = 0; k < cells.Length; k++)
                if (cells[k] != null) foreach (Point3D q in cells[k]) { mx[p] = q.motor_x; my[p] = q.motor_y; p++; }
            int[] cid = new int[n]; byte[] key = new byte[n], cls = new byte[n]; int[] perCell = new int[cells.Length];
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_l1_2d_cells(ctx, mx, my, n, off, cells.Length, e, minPts, cid, key, cls, perCell));
            for (int k = 0, p = 0; k < cells.Length; k++)
                if (cells[k] != null) foreach (Point3D q in cells[k])
                {
                    q.clusterId = cid[p]; q.isClassed = cls[p] != 0; if (key[p] != 0) q.isKeyPoint = true; p++;
                }
            return perCell;
        }

        // The whole blocked clustering (MainForm.getClusterFromMotor -> DoWork3 -> CompleteWork3 up to the renumbered labels,
        // FrmMain.cs:1214-1520) in one call.  Writes clusterId / isClassed of every point of rawData, returns MainForm.clusterSum.
        public int dbscanBlocked(List<Point3D> rawData, double e, int minPts, int ptsInCell, out int delSum)
        {
            int n = rawData.Count;
            double[] mx = new double[n], my = new double[n];
            for (int i = 0; i < n; i++) { mx[i] = rawData[i].motor_x; my[i] = rawData[i].motor_y; }
            int[] cid = new int[n];
            int clusterSum, rows, cols; long unassigned;
            NativeMethods.Check(ctx, NativeMethods.vpc_dbscan_blocked_ref(ctx, mx, my, n, e, minPts, ptsInCell, cid, out clusterSum, out delSum,
                                                                          out rows, out cols, out unassigned));
            for (int i = 0; i < n; i++) { rawData[i].clusterId = cid[i]; rawData[i].isClassed = cid[i] != 0; }
            pointsAmount += n;
            clusterAmount = clusterSum;
            return clusterSum;
        }

        public void Dispose() { if (ctx != IntPtr.Zero) { NativeMethods.vpc_destroy(ctx); ctx = IntPtr.Zero; } }
    }
}
